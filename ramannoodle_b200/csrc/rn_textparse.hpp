// Text-ingest helpers shared by the trajectory readers (rn_ingest.cu, rn_vasprun.cu): a read-only
// mmap of the file and a correctly rounded decimal parser (every value equals Python's float(token)).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "rn_common.cuh"

namespace rn {

struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    int64_t mtime_ns = 0;
    ~MappedFile() {
        if (data && size) munmap(const_cast<char*>(data), size);
        if (fd >= 0) close(fd);
    }
    int open_path(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) {
            set_error("cannot open %s: %s", path, strerror(errno));
            return RN_ERR_INVALID_ARGUMENT;
        }
        struct stat st;
        if (fstat(fd, &st) != 0) {
            set_error("cannot stat %s", path);
            return RN_ERR_INVALID_ARGUMENT;
        }
        size = (size_t)st.st_size;
        mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
        if (size == 0) {
            set_error("%s is empty", path);
            return RN_ERR_INVALID_ARGUMENT;
        }
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) {
            set_error("cannot mmap %s: %s", path, strerror(errno));
            data = nullptr;
            return RN_ERR_OUT_OF_MEMORY;
        }
        data = static_cast<const char*>(p);
        return RN_OK;
    }
};

// [begin, end) of the line starting at `pos` (end excludes the newline); returns the start of the next line
static inline size_t next_line(const char* d, size_t size, size_t pos, size_t* end) {
    const void* nl = (pos < size) ? memchr(d + pos, '\n', size - pos) : nullptr;
    if (!nl) {
        *end = size;
        return size;
    }
    *end = (size_t)(static_cast<const char*>(nl) - d);
    return *end + 1;
}

static inline const char* skip_space(const char* p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    return p;
}

// One decimal number -> double, correctly rounded like Python's float().
// Fast path (Clinger): up to 15 significant digits and |decimal exponent| <= 22 make both the
// integer mantissa and the power of ten exact doubles, so ONE multiplication or division rounds
// correctly.  Everything else (17-digit repr output, huge exponents, inf/nan) goes to glibc's
// strtod, which is correctly rounded and thread-safe.  (libstdc++'s std::from_chars<double> takes a
// process-wide lock in this toolchain and does not scale over threads.)
static inline const char* parse_double(const char* p, const char* e, double* out) {
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* start = p;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) {
        neg = (*p == '-');
        p++;
    }
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false, fast = true;
    while (p < e && *p >= '0' && *p <= '9') {
        any = true;
        if (digits < 19) {
            mant = mant * 10 + (uint64_t)(*p - '0');
            if (mant) digits++;
        } else {
            fast = false;
            exp10++;
        }
        p++;
    }
    if (p < e && *p == '.') {
        p++;
        while (p < e && *p >= '0' && *p <= '9') {
            any = true;
            if (digits < 19) {
                mant = mant * 10 + (uint64_t)(*p - '0');
                if (mant) digits++;
                exp10--;
            } else {
                fast = false;
            }
            p++;
        }
    }
    if (!any) {
        fast = false;  // inf / nan / garbage: let strtod decide
    } else if (p < e && (*p == 'e' || *p == 'E')) {
        const char* q = p + 1;
        bool eneg = false;
        if (q < e && (*q == '-' || *q == '+')) {
            eneg = (*q == '-');
            q++;
        }
        if (q < e && *q >= '0' && *q <= '9') {
            int ev = 0;
            while (q < e && *q >= '0' && *q <= '9') {
                if (ev < 10000) ev = ev * 10 + (*q - '0');
                q++;
            }
            exp10 += eneg ? -ev : ev;
            p = q;
        }
    }
    if (fast && mant <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
        double v = (double)mant;
        v = (exp10 < 0) ? v / kPow10[-exp10] : v * kPow10[exp10];
        *out = neg ? -v : v;
        return p;
    }
    // slow path: the token, NUL-terminated, through strtod
    const char* tok_end = start;
    while (tok_end < e && *tok_end != ' ' && *tok_end != '\t' && *tok_end != '\r') tok_end++;
    char buf[128];
    const size_t len = (size_t)(tok_end - start);
    if (len == 0 || len >= sizeof(buf)) return nullptr;
    memcpy(buf, start, len);
    buf[len] = 0;
    char* endp = nullptr;
    const double v = strtod(buf, &endp);
    if (endp == buf) return nullptr;
    *out = v;
    return start + (endp - buf);
}

// parse `count` whitespace-separated doubles from [p, e); returns false on failure
static inline bool parse_doubles(const char* p, const char* e, int count, double* out) {
    for (int i = 0; i < count; i++) {
        p = skip_space(p, e);
        const char* q = (p < e) ? parse_double(p, e, out + i) : nullptr;
        if (!q) return false;
        if (q < e && *q != ' ' && *q != '\t' && *q != '\r') return false;  // trailing garbage in the token
        p = q;
    }
    return true;
}

}  // namespace rn
