// Trajectory ingest: vasprun.xml molecular-dynamics runs -> (S,N,3) positions.
//
// The reference builds a full ElementTree of the file and walks it in Python
// (ramannoodle/io/vasp/vasprun.py:298-330): every `structure` element that is a DIRECT child of the
// root and carries no `name` attribute is a frame; its first `varray` child holds one `<v>` row of
// three numbers per atom (`_parse_positions`, :53-70); the timestep is the text of
// ./parameters/separator[@name='ionic']/i[@name='POTIM'] (`_parse_timestep`, :281-295).
// Here the file is mmap'ed and tokenised once by a small XML scanner (tags, attributes with quoted
// values, comments, processing instructions, CDATA, self-closing tags); the `<v>` rows of the frames
// are then converted by a pool of threads with the correctly rounded decimal parser of
// rn_textparse.hpp.  Anything the scanner does not understand (entities or markup inside a numeric row,
// rows that are not three numbers, ragged frames) is reported as RN_ERR_UNSUPPORTED, and the Python
// layer falls back to a stdlib ElementTree walk with the reference's rules.
#include <atomic>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rn_textparse.hpp"

namespace rn {

struct XmlTag {
    size_t begin = 0, end = 0;       // '<' .. one past '>'
    size_t name_begin = 0, name_end = 0;
    bool closing = false, self_closing = false;
};

// next element tag at or after `pos` (comments, processing instructions, DOCTYPE and CDATA are skipped);
// false at the end of the file; *bad is set on malformed markup
static bool next_tag(const MappedFile& f, size_t pos, XmlTag* tag, bool* bad) {
    const char* d = f.data;
    const size_t n = f.size;
    while (pos < n) {
        const void* lt = memchr(d + pos, '<', n - pos);
        if (!lt) return false;
        size_t p = (size_t)(static_cast<const char*>(lt) - d);
        if (p + 1 >= n) {
            *bad = true;
            return false;
        }
        if (d[p + 1] == '?') {  // <? ... ?>
            const void* e = memmem(d + p, n - p, "?>", 2);
            if (!e) {
                *bad = true;
                return false;
            }
            pos = (size_t)(static_cast<const char*>(e) - d) + 2;
            continue;
        }
        if (d[p + 1] == '!') {
            if (n - p >= 4 && memcmp(d + p, "<!--", 4) == 0) {
                const void* e = memmem(d + p + 4, n - p - 4, "-->", 3);
                if (!e) {
                    *bad = true;
                    return false;
                }
                pos = (size_t)(static_cast<const char*>(e) - d) + 3;
                continue;
            }
            if (n - p >= 9 && memcmp(d + p, "<![CDATA[", 9) == 0) {
                const void* e = memmem(d + p + 9, n - p - 9, "]]>", 3);
                if (!e) {
                    *bad = true;
                    return false;
                }
                pos = (size_t)(static_cast<const char*>(e) - d) + 3;
                continue;
            }
            const void* e = memchr(d + p, '>', n - p);  // <!DOCTYPE ...> (no internal subset expected)
            if (!e) {
                *bad = true;
                return false;
            }
            pos = (size_t)(static_cast<const char*>(e) - d) + 1;
            continue;
        }
        tag->begin = p;
        size_t q = p + 1;
        tag->closing = (d[q] == '/');
        if (tag->closing) q++;
        tag->name_begin = q;
        while (q < n && d[q] != '>' && d[q] != '/' && d[q] != ' ' && d[q] != '\t' && d[q] != '\n' && d[q] != '\r') q++;
        tag->name_end = q;
        if (tag->name_end == tag->name_begin) {
            *bad = true;
            return false;
        }
        // attributes: quoted values may contain '>'
        char quote = 0;
        while (q < n) {
            const char c = d[q];
            if (quote) {
                if (c == quote) quote = 0;
            } else if (c == '"' || c == '\'') {
                quote = c;
            } else if (c == '>') {
                break;
            }
            q++;
        }
        if (q >= n) {
            *bad = true;
            return false;
        }
        tag->self_closing = (!tag->closing && q > p && d[q - 1] == '/');
        tag->end = q + 1;
        return true;
    }
    return false;
}

static bool name_is(const MappedFile& f, const XmlTag& t, const char* name) {
    const size_t len = strlen(name);
    return t.name_end - t.name_begin == len && memcmp(f.data + t.name_begin, name, len) == 0;
}

// value of attribute `attr` of a start tag: [*vb, *ve); false if absent
static bool attribute(const MappedFile& f, const XmlTag& t, const char* attr, size_t* vb, size_t* ve) {
    const char* d = f.data;
    const size_t len = strlen(attr);
    size_t q = t.name_end;
    while (q < t.end) {
        while (q < t.end && (d[q] == ' ' || d[q] == '\t' || d[q] == '\n' || d[q] == '\r')) q++;
        size_t ab = q;
        while (q < t.end && d[q] != '=' && d[q] != '>' && d[q] != '/' && d[q] != ' ' && d[q] != '\t' && d[q] != '\n' &&
               d[q] != '\r')
            q++;
        const size_t ae = q;
        while (q < t.end && (d[q] == ' ' || d[q] == '\t' || d[q] == '\n' || d[q] == '\r')) q++;
        if (q >= t.end || d[q] != '=') {
            if (ae == ab) q++;
            continue;
        }
        q++;
        while (q < t.end && (d[q] == ' ' || d[q] == '\t' || d[q] == '\n' || d[q] == '\r')) q++;
        if (q >= t.end || (d[q] != '"' && d[q] != '\'')) return false;
        const char quote = d[q++];
        const size_t b = q;
        while (q < t.end && d[q] != quote) q++;
        if (q >= t.end) return false;
        if (ae - ab == len && memcmp(d + ab, attr, len) == 0) {
            *vb = b;
            *ve = q;
            return true;
        }
        q++;
    }
    return false;
}

static bool attribute_is(const MappedFile& f, const XmlTag& t, const char* attr, const char* value) {
    size_t b = 0, e = 0;
    if (!attribute(f, t, attr, &b, &e)) return false;
    return e - b == strlen(value) && memcmp(f.data + b, value, e - b) == 0;
}

struct VasprunIndex {
    std::string path;
    size_t size = 0;
    int64_t mtime_ns = 0;
    int64_t num_atoms = 0;
    double timestep = 0.0;
    bool has_timestep = false;
    std::vector<size_t> rows;  // [begin, end) text of every <v> row of every frame, frame-major: 2 entries per row
};

static std::mutex g_vasprun_mutex;
static VasprunIndex g_vasprun_index;

// One pass over the tags.  Returns RN_OK, RN_ERR_INVALID_ARGUMENT (what the reference reports as
// InvalidFileException) or RN_ERR_UNSUPPORTED (let the ElementTree fallback decide).
static int index_vasprun(const MappedFile& f, VasprunIndex& idx) {
    XmlTag tag;
    bool bad = false;
    size_t pos = 0;
    int depth = 0;             // open elements; the root is depth 1 once opened
    int frame_depth = -1;      // > 0 while inside an unnamed root-level <structure>
    int varray_depth = -1;     // > 0 while inside that structure's first <varray>
    bool frame_has_varray = false;
    int64_t rows_in_frame = 0;
    int ionic_depth = -1;      // inside ./parameters/separator[@name='ionic']
    int parameters_depth = -1;
    idx.rows.clear();
    idx.num_atoms = -1;
    idx.has_timestep = false;
    int64_t frames = 0;
    int64_t elements = 0;
    while (next_tag(f, pos, &tag, &bad)) {
        pos = tag.end;
        if (tag.closing) {
            if (depth == frame_depth) {
                if (!frame_has_varray) {
                    set_error("structure varray not found");
                    return RN_ERR_INVALID_ARGUMENT;
                }
                if (idx.num_atoms < 0) idx.num_atoms = rows_in_frame;
                if (rows_in_frame != idx.num_atoms) {
                    set_error("frames with different numbers of atoms");
                    return RN_ERR_UNSUPPORTED;
                }
                frames++;
                frame_depth = -1;
            }
            if (depth == varray_depth) varray_depth = -2;  // only the FIRST varray of a frame counts
            if (depth == ionic_depth) ionic_depth = -1;
            if (depth == parameters_depth) parameters_depth = -1;
            depth--;
            if (depth < 0) {
                bad = true;
                break;
            }
            continue;
        }
        depth++;
        elements++;
        const int this_depth = depth;
        if (this_depth == 2 && name_is(f, tag, "structure")) {
            size_t b, e;
            if (!attribute(f, tag, "name", &b, &e)) {  // named structures (initialpos, finalpos) are skipped
                frame_depth = this_depth;
                varray_depth = -1;
                frame_has_varray = false;
                rows_in_frame = 0;
                if (tag.self_closing) {
                    set_error("structure varray not found");
                    return RN_ERR_INVALID_ARGUMENT;
                }
            }
        } else if (frame_depth > 0 && this_depth == frame_depth + 1 && varray_depth == -1 && name_is(f, tag, "varray")) {
            frame_has_varray = true;
            varray_depth = tag.self_closing ? -2 : this_depth;
        } else if (varray_depth > 0 && this_depth == varray_depth + 1) {
            // every child of the varray is a row, whatever its name (the reference iterates over children)
            if (tag.self_closing) {
                set_error("varray child text not found");
                return RN_ERR_INVALID_ARGUMENT;
            }
            XmlTag close;
            bool bad2 = false;
            if (!next_tag(f, tag.end, &close, &bad2) || !close.closing) {
                set_error("markup inside a positions row");
                return RN_ERR_UNSUPPORTED;
            }
            idx.rows.push_back(tag.end);
            idx.rows.push_back(close.begin);
            rows_in_frame++;
        } else if (this_depth == 2 && name_is(f, tag, "parameters") && parameters_depth < 0 && !idx.has_timestep) {
            parameters_depth = tag.self_closing ? -1 : this_depth;
        } else if (parameters_depth > 0 && this_depth == parameters_depth + 1 && name_is(f, tag, "separator") &&
                   attribute_is(f, tag, "name", "ionic") && ionic_depth < 0 && !idx.has_timestep) {
            ionic_depth = tag.self_closing ? -1 : this_depth;
        } else if (ionic_depth > 0 && this_depth == ionic_depth + 1 && name_is(f, tag, "i") &&
                   attribute_is(f, tag, "name", "POTIM") && !idx.has_timestep) {
            if (tag.self_closing) {
                set_error("potim element has no text");
                return RN_ERR_INVALID_ARGUMENT;
            }
            XmlTag close;
            bool bad2 = false;
            if (!next_tag(f, tag.end, &close, &bad2) || !close.closing) return RN_ERR_UNSUPPORTED;
            const char* p = skip_space(f.data + tag.end, f.data + close.begin);
            while (p < f.data + close.begin && *p == '\n') p = skip_space(p + 1, f.data + close.begin);
            double v = 0.0;
            const char* q = parse_double(p, f.data + close.begin, &v);
            if (!q) {
                set_error("timestep could not be parsed");
                return RN_ERR_UNSUPPORTED;
            }
            idx.timestep = v;
            idx.has_timestep = true;
        }
        if (tag.self_closing) {
            depth--;
        }
    }
    if (bad || depth != 0 || elements == 0) {  // no root element: not an XML document
        set_error("root xml element could not be found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    if (frames == 0) {
        set_error("no trajectory found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    if (!idx.has_timestep) {
        set_error("timestep not found");
        return RN_ERR_INVALID_ARGUMENT;
    }
    return RN_OK;
}

static int cached_vasprun_index(const char* path, const MappedFile& f, VasprunIndex& out) {
    {
        std::lock_guard<std::mutex> lock(g_vasprun_mutex);
        if (g_vasprun_index.path == path && g_vasprun_index.size == f.size && g_vasprun_index.mtime_ns == f.mtime_ns &&
            !g_vasprun_index.rows.empty()) {
            out = g_vasprun_index;
            return RN_OK;
        }
    }
    int rc = index_vasprun(f, out);
    if (rc != RN_OK) return rc;
    out.path = path;
    out.size = f.size;
    out.mtime_ns = f.mtime_ns;
    std::lock_guard<std::mutex> lock(g_vasprun_mutex);
    g_vasprun_index = out;
    return RN_OK;
}

}  // namespace rn

using namespace rn;

// ramannoodle/io/vasp/vasprun.py:298-330 (read_trajectory): frame count, atom count and timestep (fs).
extern "C" int rn_vasprun_scan(const char* path, int64_t* num_frames, int64_t* num_atoms, double* timestep_fs) {
    RN_CHECK_ARG(path && num_frames && num_atoms && timestep_fs, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    VasprunIndex idx;
    rc = cached_vasprun_index(path, f, idx);
    if (rc != RN_OK) return rc;
    *num_atoms = idx.num_atoms;
    *num_frames = idx.num_atoms > 0 ? (int64_t)(idx.rows.size() / 2) / idx.num_atoms : 0;
    if (idx.num_atoms == 0) {
        set_error("frames without atoms");
        return RN_ERR_UNSUPPORTED;
    }
    *timestep_fs = idx.timestep;
    return RN_OK;
}

// Fills h_positions (num_frames, num_atoms, 3) with the fractional coordinates as written, or — wrap != 0 —
// wrapped into [0,1) as x - floor(x) (Trajectory.__init__, dynamics/_trajectory.py:45).
extern "C" int rn_vasprun_read(const char* path, double* h_positions, int64_t num_frames, int64_t num_atoms,
                               int num_threads, int wrap) {
    RN_CHECK_ARG(path && h_positions, "null pointer");
    MappedFile f;
    int rc = f.open_path(path);
    if (rc != RN_OK) return rc;
    VasprunIndex idx;
    rc = cached_vasprun_index(path, f, idx);
    if (rc != RN_OK) return rc;
    const int64_t rows = (int64_t)(idx.rows.size() / 2);
    RN_CHECK_ARG(idx.num_atoms == num_atoms && rows == num_frames * num_atoms,
                 "file holds %lld rows of %lld atoms, buffer was sized for %lld x %lld", (long long)rows,
                 (long long)idx.num_atoms, (long long)num_frames, (long long)num_atoms);
    if (num_threads <= 0) num_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    num_threads = (int)std::min<int64_t>(num_threads, std::max<int64_t>(1, rows / 4096 + 1));
    std::atomic<int64_t> bad_row{-1};
    auto convert = [&](int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1 && bad_row.load(std::memory_order_relaxed) < 0; r++) {
            const char* b = f.data + idx.rows[(size_t)(2 * r)];
            const char* e = f.data + idx.rows[(size_t)(2 * r + 1)];
            while (e > b && (e[-1] == ' ' || e[-1] == '\n' || e[-1] == '\r' || e[-1] == '\t')) e--;
            while (b < e && (*b == '\n' || *b == ' ' || *b == '\t' || *b == '\r')) b++;
            double* out = h_positions + r * 3;
            // exactly three numbers and nothing else (text.split() of the reference yields three tokens)
            if (memchr(b, '\n', (size_t)(e - b)) != nullptr || memchr(b, '&', (size_t)(e - b)) != nullptr ||
                !parse_doubles(b, e, 3, out) ) {
                bad_row.store(r);
                return;
            }
            // parse_doubles stops after the third token: anything left in the row is unexpected
            {
                const char* p = b;
                int tokens = 0;
                while (p < e) {
                    p = skip_space(p, e);
                    if (p >= e) break;
                    tokens++;
                    while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
                }
                if (tokens != 3) {
                    bad_row.store(r);
                    return;
                }
            }
            if (wrap)
                for (int c = 0; c < 3; c++) out[c] -= floor(out[c]);
        }
    };
    if (num_threads == 1) {
        convert(0, rows);
    } else {
        const int64_t block = 4096;
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const int64_t r0 = next.fetch_add(block);
                if (r0 >= rows) return;
                convert(r0, std::min<int64_t>(rows, r0 + block));
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < num_threads; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    if (bad_row.load() >= 0) {
        set_error("positions row %lld is not three numbers", (long long)bad_row.load());
        return RN_ERR_UNSUPPORTED;
    }
    return RN_OK;
}
