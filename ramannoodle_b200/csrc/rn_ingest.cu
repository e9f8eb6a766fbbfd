// Trajectory ingest: VASP XDATCAR -> (S,N,3) fractional positions.
//
// The reference parses XDATCAR files line by line in Python (ramannoodle/io/vasp/xdatcar.py:21-56
// through io/vasp/poscar.py:_read_lattice/_read_atomic_symbols/_read_positions); once the GPU
// evaluates a million frames in under a millisecond, that text parsing is the end-to-end wall.
// This is a host-side C++ reader: the file is mmap'ed, frame offsets are found in one pass over
// the newlines, and frames are converted by a pool of threads with a correctly rounded decimal parser
// (exact fast path + strtod fallback, matching Python's float()), straight into the caller's buffer
// (e.g. pinned memory).
// Semantics follow the reference reader: title line, scale factor, three lattice rows, symbols,
// counts, then frames of one label line (first character D/d = direct) + N coordinate lines; the
// series ends at the first blank/missing label line.  Cartesian frames are not supported here.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rn_textparse.hpp"

namespace rn {

struct XdatcarHeader {
    double lattice[9];
    int64_t num_atoms = 0;
    size_t frames_begin = 0;  // byte offset of the first frame's label line
};

static int parse_header(const MappedFile& f, XdatcarHeader& h) {
    size_t pos = 0, end = 0;
    pos = next_line(f.data, f.size, pos, &end);  // title
    size_t line = pos;
    pos = next_line(f.data, f.size, pos, &end);
    double scale = 0;
    if (!parse_doubles(f.data + line, f.data + end, 1, &scale)) {
        set_error("scale factor could not be parsed");
        return RN_ERR_INVALID_ARGUMENT;
    }
    for (int r = 0; r < 3; r++) {
        line = pos;
        pos = next_line(f.data, f.size, pos, &end);
        if (!parse_doubles(f.data + line, f.data + end, 3, h.lattice + 3 * r)) {
            set_error("lattice could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
        for (int c = 0; c < 3; c++) h.lattice[3 * r + c] *= scale;
    }
    // symbols line: count tokens
    line = pos;
    pos = next_line(f.data, f.size, pos, &end);
    int symbols = 0;
    for (const char* p = skip_space(f.data + line, f.data + end); p < f.data + end; p = skip_space(p, f.data + end)) {
        symbols++;
        while (p < f.data + end && *p != ' ' && *p != '\t' && *p != '\r') p++;
    }
    if (symbols == 0) {
        set_error("no atom symbols found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    // counts line
    line = pos;
    pos = next_line(f.data, f.size, pos, &end);
    int counts = 0;
    int64_t atoms = 0;
    for (const char* p = skip_space(f.data + line, f.data + end); p < f.data + end; p = skip_space(p, f.data + end)) {
        int64_t v = 0;
        auto res = std::from_chars(p, f.data + end, v);
        if (res.ec != std::errc() || v < 0) {
            set_error("could not parse ion counts");
            return RN_ERR_INVALID_ARGUMENT;
        }
        atoms += v;
        counts++;
        p = res.ptr;
    }
    if (counts != symbols) {
        set_error("wrong number of ion counts: %d != %d", counts, symbols);
        return RN_ERR_INVALID_ARGUMENT;
    }
    if (atoms <= 0) {
        set_error("no atoms in file");
        return RN_ERR_INVALID_ARGUMENT;
    }
    h.num_atoms = atoms;
    h.frames_begin = pos;
    return RN_OK;
}

// offsets of the first coordinate line of every frame
static int scan_frames(const MappedFile& f, const XdatcarHeader& h, std::vector<size_t>& starts) {
    size_t pos = h.frames_begin, end = 0;
    while (pos < f.size) {
        const size_t label = pos;
        pos = next_line(f.data, f.size, pos, &end);
        const char* p = f.data + label;
        const char* e = f.data + end;
        // the reference stops at a label line that is empty or starts with whitespace
        if (p >= e || *p == ' ' || *p == '\t' || *p == '\r') break;
        char c = *p;
        if (c == 's' || c == 'S') {  // selective dynamics: the coordinate format is on the next line
            const size_t l2 = pos;
            pos = next_line(f.data, f.size, pos, &end);
            if (l2 >= f.size) break;
            c = f.data[l2];
        }
        if (c == 'c' || c == 'C') {
            set_error("Cartesian XDATCAR frames are not supported by the fast reader");
            return RN_ERR_UNSUPPORTED;
        }
        if (c != 'd' && c != 'D') {
            set_error("unrecognized coordinate format in frame %zu", starts.size() + 1);
            return RN_ERR_INVALID_ARGUMENT;
        }
        starts.push_back(pos);
        for (int64_t a = 0; a < h.num_atoms; a++) {
            if (pos >= f.size) {
                set_error("positions could not be parsed: file ends inside frame %zu", starts.size());
                return RN_ERR_INVALID_ARGUMENT;
            }
            pos = next_line(f.data, f.size, pos, &end);
        }
    }
    return RN_OK;
}

// ---- OUTCAR molecular-dynamics trajectories (ramannoodle/io/vasp/outcar.py:497-538) ----------
//
// The reference walks the file once, front to back: POTCAR block -> "ions per type" (atom count,
// :46-86), "time-step for ionic-motion" (:481-494), "Write flags" then "direct lattice vectors"
// (:212-241), then every "POSITION ... TOTAL-FORCE" block: one separator line and N lines whose
// first three numbers are Cartesian coordinates.  With machine-learned force fields an ab-initio
// block that directly follows an "(ML)" block repeats the same step and is skipped (:520-527).

struct OutcarHeader {
    int64_t num_atoms = 0;
    double timestep = 0.0;
    double lattice[9];
    size_t frames_begin = 0;
};

// first line at or after `pos` that contains `needle`: [line, end) and the start of the next line
static bool find_line(const MappedFile& f, size_t pos, const char* needle, size_t* line, size_t* end, size_t* next) {
    if (pos >= f.size) return false;
    const void* hit = memmem(f.data + pos, f.size - pos, needle, strlen(needle));
    if (!hit) return false;
    size_t at = (size_t)(static_cast<const char*>(hit) - f.data);
    size_t b = at;
    while (b > pos && f.data[b - 1] != '\n') b--;
    *line = b;
    *next = next_line(f.data, f.size, b, end);
    return true;
}

static bool line_contains(const MappedFile& f, size_t line, size_t end, const char* needle) {
    return memmem(f.data + line, end - line, needle, strlen(needle)) != nullptr;
}

static const char* const kElementSymbols[] = {
    "H",  "He", "Li", "Be", "B",  "C",  "N",  "O",  "F",  "Ne", "Na", "Mg", "Al", "Si", "P",  "S",  "Cl", "Ar", "K",  "Ca",
    "Sc", "Ti", "V",  "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",  "Zr",
    "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I",  "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd",
    "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W",  "Re", "Os", "Ir", "Pt", "Au", "Hg",
    "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U",  "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm",
    "Md", "No", "Lr", "Rf", "Db", "Sg", "Bh", "Hs", "Mt", "Ds", "Rg", "Cn", "Nh", "Fl", "Mc", "Lv", "Ts", "Og"};

// outcar.py:27-43: second-to-last token of a POTCAR line, cut at '_', must be an element symbol
static bool potcar_symbol_ok(const MappedFile& f, size_t line, size_t end) {
    const char* toks[64];
    size_t lens[64];
    int count = 0;
    const char* p = f.data + line;
    const char* e = f.data + end;
    while (p < e) {
        p = skip_space(p, e);
        if (p >= e) break;
        const char* b = p;
        while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
        if (count < 64) {
            toks[count] = b;
            lens[count] = (size_t)(p - b);
            count++;
        }
    }
    if (count < 2) return false;
    const char* tok = toks[count - 2];
    size_t len = lens[count - 2];
    for (size_t i = 0; i < len; i++)
        if (tok[i] == '_') len = i;
    for (const char* sym : kElementSymbols)
        if (strlen(sym) == len && memcmp(sym, tok, len) == 0) return true;
    return false;
}

static int parse_outcar_header(const MappedFile& f, OutcarHeader& h) {
    size_t line = 0, end = 0, pos = 0;
    if (!find_line(f, 0, "POTCAR:    ", &line, &end, &pos)) {
        set_error("POTCAR block not found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    int symbols = 0;
    for (;;) {
        if (!potcar_symbol_ok(f, line, end)) {
            set_error("POTCAR block could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
        symbols++;
        if (pos >= f.size) break;
        line = pos;
        pos = next_line(f.data, f.size, pos, &end);
        if (!line_contains(f, line, end, "POTCAR")) break;
    }
    if (pos < f.size) {  // the line after the block: a VRHFIN line means the block listed one entry twice
        line = pos;
        pos = next_line(f.data, f.size, pos, &end);
        if (line_contains(f, line, end, "VRHFIN")) symbols--;
    }
    if (!find_line(f, pos, "ions per type", &line, &end, &pos)) {
        set_error("ion number block could not be parsed");
        return RN_ERR_INVALID_ARGUMENT;
    }
    {
        int token = 0, used = 0;
        int64_t atoms = 0;
        const char* p = f.data + line;
        const char* e = f.data + end;
        while (p < e) {
            p = skip_space(p, e);
            if (p >= e) break;
            const char* b = p;
            while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
            if (token >= 4) {
                int64_t v = 0;
                auto res = std::from_chars(b, p, v);
                if (res.ec != std::errc() || res.ptr != p) {
                    set_error("ion number block could not be parsed");
                    return RN_ERR_INVALID_ARGUMENT;
                }
                if (used < symbols) atoms += v;  // zip(potcar_symbols, atomic_numbers)
                used++;
            }
            token++;
        }
        if (atoms <= 0) {
            set_error("ion number block could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
        h.num_atoms = atoms;
    }
    if (!find_line(f, pos, "time-step for ionic-motion", &line, &end, &pos)) {
        set_error("timestep not found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    {
        const char* p = f.data + line;
        const char* e = f.data + end;
        const char* tok = nullptr;
        const char* tok_end = nullptr;
        for (int t = 0; t < 3; t++) {
            p = skip_space(p, e);
            tok = p;
            while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
            tok_end = p;
        }
        double v = 0;
        if (!tok || tok == tok_end || parse_double(tok, tok_end, &v) != tok_end) {
            set_error("timestep could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
        h.timestep = v;
    }
    if (!find_line(f, pos, "Write flags", &line, &end, &pos) ||
        !find_line(f, pos, "direct lattice vectors      ", &line, &end, &pos)) {
        set_error("outcar does not have expected format");
        return RN_ERR_INVALID_ARGUMENT;
    }
    for (int r = 0; r < 3; r++) {
        if (pos >= f.size) {
            set_error("lattice could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
        line = pos;
        pos = next_line(f.data, f.size, pos, &end);
        if (!parse_doubles(f.data + line, f.data + end, 3, h.lattice + 3 * r)) {
            set_error("lattice could not be parsed");
            return RN_ERR_INVALID_ARGUMENT;
        }
    }
    h.frames_begin = pos;
    return RN_OK;
}

// offsets of the first coordinate line of every frame the reference keeps
static int scan_outcar_frames(const MappedFile& f, const OutcarHeader& h, std::vector<size_t>& starts) {
    static const char kBlock[] = "POSITION                                       TOTAL-FORCE ";
    size_t pos = h.frames_begin, line = 0, end = 0;
    bool ml_step = false;
    while (find_line(f, pos, kBlock, &line, &end, &pos)) {
        const bool is_ml = line_contains(f, line, end, "(ML)");
        if (ml_step && !is_ml) {  // ab-initio repeat of the ML step just read
            ml_step = false;
            continue;
        }
        ml_step = is_ml;
        if (pos < f.size) pos = next_line(f.data, f.size, pos, &end);  // separator line
        starts.push_back(pos);
        for (int64_t a = 0; a < h.num_atoms; a++) {
            if (pos >= f.size) {
                set_error("Cartesian positions could not be parsed: file ends inside frame %zu", starts.size());
                return RN_ERR_INVALID_ARGUMENT;
            }
            pos = next_line(f.data, f.size, pos, &end);
        }
    }
    if (starts.empty()) {
        set_error("no trajectory found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    return RN_OK;
}

// Threaded frame scan for regular files: every thread walks the lines of its byte range and notes the
// lines that start with D/d (coordinate lines start with blanks, digits or a sign, never a letter).
// The result is accepted only if it is exactly what the sequential reference walk would produce —
// first label right after the header, labels exactly N+1 lines apart, N complete lines after the last
// one, nothing but blank lines behind them; anything else (selective dynamics, Cartesian frames,
// blank or extra lines in between) is left to scan_frames, which follows the reference line by line.
static bool scan_frames_parallel(const MappedFile& f, const XdatcarHeader& h, std::vector<size_t>& starts,
                                 int num_threads) {
    const size_t begin = h.frames_begin;
    if (begin >= f.size || num_threads < 2) return false;
    const size_t span = f.size - begin;
    if (span < ((size_t)8 << 20)) return false;
    num_threads = (int)std::min<size_t>((size_t)num_threads, span >> 20);
    struct Chunk {
        std::vector<size_t> first_coord;  // offset of the line after each label line
        std::vector<int64_t> label_line;  // line index (within the chunk) of each label line
        int64_t lines = 0;                // lines starting in this chunk
        bool odd = false;                 // a line starting with a letter other than D/d
    };
    std::vector<Chunk> chunks((size_t)num_threads);
    auto work = [&](int t) {
        Chunk& c = chunks[(size_t)t];
        size_t lo = begin + span * (size_t)t / (size_t)num_threads;
        const size_t hi = begin + span * (size_t)(t + 1) / (size_t)num_threads;
        if (t > 0) {  // first line that starts at or after lo
            const void* nl = memchr(f.data + lo - 1, '\n', f.size - (lo - 1));
            if (!nl) return;
            lo = (size_t)(static_cast<const char*>(nl) - f.data) + 1;
        }
        size_t pos = lo, end = 0;
        while (pos < hi && pos < f.size) {
            const char ch = f.data[pos];
            const size_t next = next_line(f.data, f.size, pos, &end);
            if (ch == 'D' || ch == 'd') {
                c.first_coord.push_back(next);
                c.label_line.push_back(c.lines);
            } else if ((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z')) {
                c.odd = true;
            }
            c.lines++;
            pos = next;
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < num_threads; t++) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    }
    // stitch and verify
    int64_t base = 0, prev_label = -1, total_lines = 0;
    size_t count = 0;
    for (const Chunk& c : chunks) {
        if (c.odd) return false;
        count += c.label_line.size();
        total_lines += c.lines;
    }
    if (count == 0) return false;
    starts.clear();
    starts.reserve(count);
    for (const Chunk& c : chunks) {
        for (size_t i = 0; i < c.label_line.size(); i++) {
            const int64_t line = base + c.label_line[i];
            if (prev_label < 0 ? line != 0 : line != prev_label + h.num_atoms + 1) {
                starts.clear();
                return false;
            }
            prev_label = line;
            starts.push_back(c.first_coord[i]);
        }
        base += c.lines;
    }
    // the last frame must be complete; whatever follows it must be blank lines (the reference stops there)
    if (total_lines < prev_label + 1 + h.num_atoms) {
        starts.clear();
        return false;
    }
    size_t pos = starts.back(), end = 0;
    for (int64_t a = 0; a < h.num_atoms; a++) pos = next_line(f.data, f.size, pos, &end);
    while (pos < f.size) {
        const size_t line = pos;
        pos = next_line(f.data, f.size, pos, &end);
        for (size_t i = line; i < end; i++)
            if (f.data[i] != ' ' && f.data[i] != '\t' && f.data[i] != '\r') {
                starts.clear();
                return false;
            }
    }
    return true;
}

// rn_xdatcar_scan is always followed by rn_xdatcar_read on the same file: keep the last frame index
// so the newline pass runs once (keyed by path, size and mtime).
struct FrameIndex {
    std::string path;
    size_t size = 0;
    int64_t mtime_ns = 0;
    XdatcarHeader header;
    std::vector<size_t> starts;
};
static std::mutex g_index_mutex;
static FrameIndex g_index;

static int index_file(const char* path, const MappedFile& f, XdatcarHeader& h, std::vector<size_t>& starts,
                      bool consume) {
    {
        std::lock_guard<std::mutex> lock(g_index_mutex);
        if (g_index.path == path && g_index.size == f.size && g_index.mtime_ns == f.mtime_ns &&
            !g_index.starts.empty()) {
            h = g_index.header;
            if (consume) {
                starts.swap(g_index.starts);
                g_index.path.clear();
            } else {
                starts = g_index.starts;
            }
            return RN_OK;
        }
    }
    int rc = parse_header(f, h);
    if (rc != RN_OK) return rc;
    if (!scan_frames_parallel(f, h, starts, (int)std::max(1u, std::thread::hardware_concurrency()))) {
        starts.clear();
        rc = scan_frames(f, h, starts);
        if (rc != RN_OK) return rc;
    }
    if (!consume) {
        std::lock_guard<std::mutex> lock(g_index_mutex);
        g_index.path = path;
        g_index.size = f.size;
        g_index.mtime_ns = f.mtime_ns;
        g_index.header = h;
        g_index.starts = starts;
    }
    return RN_OK;
}

}  // namespace rn

using namespace rn;

// ramannoodle/io/vasp/xdatcar.py:21-56 (header + frame count); lattice rows are scaled lattice vectors.
extern "C" int rn_xdatcar_scan(const char* path, int64_t* num_frames, int64_t* num_atoms, double* lattice) {
    RN_CHECK_ARG(path && num_frames && num_atoms, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    XdatcarHeader h;
    std::vector<size_t> starts;
    rc = index_file(path, f, h, starts, false);
    if (rc != RN_OK) return rc;
    *num_frames = (int64_t)starts.size();
    *num_atoms = h.num_atoms;
    if (lattice) memcpy(lattice, h.lattice, sizeof(h.lattice));
    return RN_OK;
}

// Fills h_positions (num_frames, num_atoms, 3) with the fractional coordinates as written, or —
// wrap != 0 — wrapped into [0,1) as x - floor(x), which is what Trajectory.__init__ applies
// (ramannoodle/dynamics/trajectory.py:58, structure/structure_utils.py:apply_pbc).
// num_frames / num_atoms must match rn_xdatcar_scan.  num_threads <= 0: all cores.
extern "C" int rn_xdatcar_read(const char* path, double* h_positions, int64_t num_frames, int64_t num_atoms,
                               int num_threads, int wrap) {
    RN_CHECK_ARG(path && h_positions, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    XdatcarHeader h;
    std::vector<size_t> starts;
    rc = index_file(path, f, h, starts, true);
    if (rc != RN_OK) return rc;
    RN_CHECK_ARG((int64_t)starts.size() == num_frames && h.num_atoms == num_atoms,
                 "file holds %zu frames of %lld atoms, buffer was sized for %lld x %lld", starts.size(),
                 (long long)h.num_atoms, (long long)num_frames, (long long)num_atoms);
    if (num_threads <= 0) num_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    num_threads = (int)std::min<int64_t>(num_threads, std::max<int64_t>(1, num_frames));
    std::atomic<int64_t> bad_frame{-1};
    auto convert = [&](int64_t f0, int64_t f1) {
        for (int64_t fr = f0; fr < f1 && bad_frame.load(std::memory_order_relaxed) < 0; fr++) {
            size_t pos = starts[(size_t)fr], end = 0;
            double* out = h_positions + fr * num_atoms * 3;
            for (int64_t a = 0; a < num_atoms; a++) {
                const size_t line = pos;
                pos = next_line(f.data, f.size, pos, &end);
                if (!parse_doubles(f.data + line, f.data + end, 3, out + 3 * a)) {
                    bad_frame.store(fr);
                    return;
                }
                if (wrap) {
                    for (int c = 0; c < 3; c++) out[3 * a + c] -= floor(out[3 * a + c]);
                }
            }
        }
    };
    if (num_threads == 1) {
        convert(0, num_frames);
    } else {
        // frames are handed out in small blocks so a descheduled thread does not hold up the rest
        const int64_t block = std::max<int64_t>(1, std::min<int64_t>(64, num_frames / (8 * num_threads)));
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const int64_t f0 = next.fetch_add(block);
                if (f0 >= num_frames) return;
                convert(f0, std::min<int64_t>(num_frames, f0 + block));
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < num_threads; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    if (bad_frame.load() >= 0) {
        set_error("positions could not be parsed in frame %lld", (long long)bad_frame.load() + 1);
        return RN_ERR_INVALID_ARGUMENT;
    }
    return RN_OK;
}

// ramannoodle/io/vasp/outcar.py:497-538 (read_trajectory): frame count, atom count, lattice (rows
// are lattice vectors, Angstrom) and the timestep (fs) of an OUTCAR molecular-dynamics run.
extern "C" int rn_outcar_scan(const char* path, int64_t* num_frames, int64_t* num_atoms, double* lattice,
                              double* timestep_fs) {
    RN_CHECK_ARG(path && num_frames && num_atoms, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    OutcarHeader h;
    rc = parse_outcar_header(f, h);
    if (rc != RN_OK) return rc;
    std::vector<size_t> starts;
    rc = scan_outcar_frames(f, h, starts);
    if (rc != RN_OK) return rc;
    *num_frames = (int64_t)starts.size();
    *num_atoms = h.num_atoms;
    if (lattice) memcpy(lattice, h.lattice, sizeof(h.lattice));
    if (timestep_fs) *timestep_fs = h.timestep;
    return RN_OK;
}

// Fills h_positions (num_frames, num_atoms, 3).  inv_lattice == NULL: the Cartesian coordinates as
// written.  Otherwise fractional coordinates cart @ inv_lattice (row-major 3x3, what
// outcar.py:529 computes per frame), wrapped into [0,1) when wrap != 0.
extern "C" int rn_outcar_read(const char* path, double* h_positions, int64_t num_frames, int64_t num_atoms,
                              const double* inv_lattice, int num_threads, int wrap) {
    RN_CHECK_ARG(path && h_positions, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    OutcarHeader h;
    rc = parse_outcar_header(f, h);
    if (rc != RN_OK) return rc;
    std::vector<size_t> starts;
    rc = scan_outcar_frames(f, h, starts);
    if (rc != RN_OK) return rc;
    RN_CHECK_ARG((int64_t)starts.size() == num_frames && h.num_atoms == num_atoms,
                 "file holds %zu frames of %lld atoms, buffer was sized for %lld x %lld", starts.size(),
                 (long long)h.num_atoms, (long long)num_frames, (long long)num_atoms);
    if (num_threads <= 0) num_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    num_threads = (int)std::min<int64_t>(num_threads, std::max<int64_t>(1, num_frames));
    std::atomic<int64_t> bad_frame{-1};
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t fr = next.fetch_add(1);
            if (fr >= num_frames || bad_frame.load(std::memory_order_relaxed) >= 0) return;
            size_t pos = starts[(size_t)fr], end = 0;
            double* out = h_positions + fr * num_atoms * 3;
            for (int64_t a = 0; a < num_atoms; a++) {
                const size_t line = pos;
                pos = next_line(f.data, f.size, pos, &end);
                double c[3];
                if (!parse_doubles(f.data + line, f.data + end, 3, c)) {
                    bad_frame.store(fr);
                    return;
                }
                for (int j = 0; j < 3; j++) {
                    double v = c[j];
                    if (inv_lattice) v = c[0] * inv_lattice[j] + c[1] * inv_lattice[3 + j] + c[2] * inv_lattice[6 + j];
                    if (inv_lattice && wrap) v -= floor(v);
                    out[3 * a + j] = v;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < num_threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (bad_frame.load() >= 0) {
        set_error("Cartesian positions could not be parsed in frame %lld", (long long)bad_frame.load() + 1);
        return RN_ERR_INVALID_ARGUMENT;
    }
    return RN_OK;
}

// Trajectory.__init__ on a host array (dynamics/_trajectory.py:45 -> structure/utils.py:27:
// positions - positions // 1), threaded: numpy's floor_divide runs at ~0.3 GB/s on one core, which
// makes wrapping a 1M-frame trajectory (4.6 GB) cost far more than its evaluation on the GPU.
// out[i] = in[i] - floor(in[i]) (identical to numpy for every finite input, -0.0, NaN and Inf);
// out may be page-locked memory, in == out is allowed.  num_threads <= 0: all cores.
extern "C" int rn_host_apply_pbc(const double* h_in, double* h_out, int64_t count, int num_threads) {
    RN_CHECK_ARG(count >= 0, "count must be non-negative");
    if (count == 0) return RN_OK;
    RN_CHECK_ARG(h_in && h_out, "null pointer");
    if (num_threads <= 0) num_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t block = 1 << 18;  // 2 MiB of doubles per work item
    num_threads = (int)std::min<int64_t>(num_threads, (count + block - 1) / block);
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t b0 = next.fetch_add(block);
            if (b0 >= count) return;
            const int64_t b1 = std::min(count, b0 + block);
            for (int64_t i = b0; i < b1; i++) h_out[i] = h_in[i] - floor(h_in[i]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < num_threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return RN_OK;
}
