// MD Raman spectrum kernels (sm_100a): MDRamanSpectrum.measure
// (ramannoodle/spectrum/_raman.py:241-309) and calc_signal_spectrum
// (ramannoodle/spectrum/utils.py:76-124).
//
// The reference forms the linear autocorrelation of each signal x (length M = S-1) with
// scipy.signal.correlate and takes the real part of its length-M FFT.  We use the exact
// identity  Re FFT_M(ac+)[k] = (|FFT_M(x)[k]|^2 + sum_n x_n^2) / 2  (SURVEY.md §7 step 7), so the
// only transform needed is FFT_M(x) for arbitrary M.  That is computed with Bluestein's
// chirp-z algorithm on top of a hand-written power-of-two Stockham FFT (radix 8/4/2 passes,
// twiddles from sincospi).  The six distinct tensor components are packed pairwise into
// three complex sequences; the invariant combinations (trace, xx-yy, ...) are formed in the
// frequency domain (FFT linearity), their energies in the time domain.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "rn_common.cuh"

struct rn_spectrum_plan {
    int device = 0;
    int sm_count = 0;
    int64_t S = 0, M = 0, L = 0;
    int log2l = 0;
    int log2tile = 10;
    int num_passes = 0;
    int pass_log2r[4] = {0, 0, 0, 0};
    int split = 0;                // two-level twiddle split: m = hi << split | lo
    double2* d_buf0 = nullptr;    // L
    double2* d_buf1 = nullptr;    // L
    double2* d_filter = nullptr;  // L   FFT of the chirp filter
    double2* d_spec = nullptr;    // 3*M  chirp-z outputs (unscaled by 1/L)
    double2* d_chirp = nullptr;   // M   exp(-i pi n^2 / M)
    double2* d_whi = nullptr;     // L >> split   exp(-2 pi i (a << split) / L)
    double2* d_wlo = nullptr;     // 1 << split   exp(-2 pi i b / L)
    double2* d_wsub = nullptr;    // 4096         exp(-2 pi i m / 4096)
    double* d_partial = nullptr;  // energy_blocks * 8
    double* d_energy = nullptr;   // 8
    int energy_blocks = 0;
    // half-length machinery for transforms split over two ranks (rn_md_spectrum_half): a plan of
    // length L/2 that shares the chirp table, and the filter spectrum decimated by residue
    rn_spectrum_plan* half = nullptr;
    double2* d_hhalf[2] = {nullptr, nullptr};  // H[r + 2k'], k' < L/2
    bool owns_chirp = true;
    // 0: the part / half entries compute the series energies themselves; 1: they treat them as zero and
    // the caller adds the constant of rn_series_energy_constant (sharded over ranks) to every bin
    int energy_mode = 0;
};

namespace rn {


__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// multiply by sgn*i
template <int SGN>
__device__ __forceinline__ double2 mul_i(double2 a) {
    return SGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// exp(-i*pi*n^2/M) (forward chirp); n^2 mod 2M is formed exactly in 64-bit integers
__device__ __forceinline__ double2 chirp(int64_t n, int64_t M) {
    const uint64_t m = ((uint64_t)n * (uint64_t)n) % (uint64_t)(2 * M);
    double s, c;
    sincospi((double)m / (double)M, &s, &c);
    return make_double2(c, -s);
}

template <int SGN>
__device__ __forceinline__ void dft2(double2* v) {
    const double2 a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
}

template <int SGN>
__device__ __forceinline__ void dft4(double2* v) {
    const double2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    const double2 s13 = cadd(v[1], v[3]), d13 = mul_i<SGN>(csub(v[1], v[3]));
    v[0] = cadd(s02, s13);
    v[1] = cadd(d02, d13);
    v[2] = csub(s02, s13);
    v[3] = csub(d02, d13);
}

template <int SGN>
__device__ __forceinline__ void dft8(double2* v) {
    double2 e[4] = {v[0], v[2], v[4], v[6]};
    double2 o[4] = {v[1], v[3], v[5], v[7]};
    dft4<SGN>(e);
    dft4<SGN>(o);
    const double h = 0.70710678118654752440;
    // W8^1 = (1 + sgn*i)/sqrt2, W8^2 = sgn*i, W8^3 = (-1 + sgn*i)/sqrt2
    const double2 o1 = make_double2(h * (o[1].x - SGN * o[1].y), h * (SGN * o[1].x + o[1].y));
    const double2 o2 = mul_i<SGN>(o[2]);
    const double2 o3 = make_double2(h * (-o[3].x - SGN * o[3].y), h * (SGN * o[3].x - o[3].y));
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);
    v[7] = csub(e[3], o3);
}

// np.diff of the series (_raman.py:282) packed two real signals per complex sequence: (c1, c2) >= 0
// are tensor components; c1 == -1: (xx - yy, yy - zz), the anisotropy differences (_raman.py:289-291);
// c1 == -2: (trace, xy) (_raman.py:286-288,292)
__device__ __forceinline__ void alpha_signal(const double* __restrict__ src, int64_t n, int c1, int c2, double& d1,
                                             double& d2) {
    const double* a = src + n * 9;
    if (c1 >= 0) {
        d1 = __ldg(a + 9 + c1) - __ldg(a + c1);
        d2 = __ldg(a + 9 + c2) - __ldg(a + c2);
    } else {
        const double xx = __ldg(a + 9) - __ldg(a), yy = __ldg(a + 13) - __ldg(a + 4);
        const double zz = __ldg(a + 17) - __ldg(a + 8);
        if (c1 == -1) {
            d1 = xx - yy;
            d2 = yy - zz;
        } else {
            d1 = xx + yy + zz;
            d2 = __ldg(a + 10) - __ldg(a + 1);
        }
    }
}

// ---- tiled Stockham FFT ---------------------------------------------------------------------
// A length-L (power of two) transform is 1-4 global passes.  Pass i is a radix-R_i Stockham step
// (R_i <= 256): out[(j-k) R + k + y Ns] = DFT_R( in[j + x L/R] * W_{Ns R}^{k x} )[y], k = j mod Ns.
// A CTA tile is B consecutive j (B*R = 1024 elements, 8 per thread, 128 threads): the tile is
// gathered in B-element contiguous chunks, the R-point DFTs run cooperatively in shared memory
// as radix-8/4/2 Stockham sub-passes, and the result is scattered in contiguous chunks.  Many
// small CTAs are resident per SM, so one CTA's gather latency and barriers overlap another's
// butterflies (the transform is FP64-issue/latency bound on B200, not HBM bound: ~120 FP64
// instructions per element per transform against a 64-lane/clk FP64 pipe).  Twiddles come from exact (80-bit,
// host-computed) tables: a two-level table for W_L and a 4096-entry table for the sub-pass
// twiddles.  The Bluestein pre-multiply (load), the filter multiply (store of the forward
// transform) and the chirp post-multiply (store of the inverse transform) are fused in.
enum { LOAD_PLAIN = 0, LOAD_ALPHA = 1, LOAD_SIGNAL = 2, LOAD_ALPHA_TW = 3 };  // _TW: times W_{2L}^n (split transforms)
enum { STORE_PLAIN = 0, STORE_POST = 1, STORE_MULH = 2 };

constexpr int kMaxLog2R = 8;  // sub-transforms of at most 256 points per pass
// CTA tile: 1024 complex elements (128 threads, many CTAs per SM) while the transform is L2 resident;
// 4096 elements (512 threads, 256-byte global chunks) for transforms that stream from HBM.
constexpr int kLog2TileSmall = 10, kLog2TileLarge = 11;
constexpr int kLargeTileMinLog2L = 22;

struct PassParams {
    int64_t L;
    int log2r;       // R = 1 << log2r
    int log2b;       // B = 1 << log2b
    int64_t Ns;      // product of the radices of earlier passes
    int64_t tw_stride;  // L / (Ns R)
    int split;
    const double2* whi;
    const double2* wlo;
    const double2* wsub;
    const double2* in;
    double2* out;
    const double2* H;      // STORE_MULH
    const double* src;     // LOAD_ALPHA / LOAD_SIGNAL
    int c1, c2;            // tensor components packed into (re, im)
    int64_t M;
    const double2* chirp;  // LOAD_ALPHA / LOAD_SIGNAL / STORE_POST
    double2* spec;         // STORE_POST
    const double2* rwhi;   // LOAD_ALPHA_TW: two-level table of the double-length transform
    const double2* rwlo;
    int rsplit;
};

// RAD: radix of the first sub-pass (8 whenever R >= 8; it carries the pass twiddle, and one radix-8
// butterfly per thread needs one two-level table lookup where radix-2 butterflies need four).
// LAST: radix of a final radix-4 / radix-2 sub-pass (8 = none) that completes R = 8^m * LAST.
template <int SGN, int LOAD, int STORE, int RAD, int LAST>
__device__ __forceinline__ void fft_tile_body(const PassParams& P, double2* S) {
    const int R = 1 << P.log2r, B = 1 << P.log2b;
    const int pitch = (B > 1) ? B + 1 : 1;  // padded batch pitch: conflict-free transposing store
    const int tile = R << P.log2b;
    const int T = tile >> 3 > 0 ? tile >> 3 : 1;  // active threads (8 elements each)
    const int tid = threadIdx.x;
    const bool active = tid < T;
    // every element index fits in 32 bits (L <= 2^30, checked at plan creation); strides are powers of two
    typedef uint32_t idx_t;
    const int log2cnt = 63 - __clzll((unsigned long long)P.L) - P.log2r;  // cnt = L / R = number of j
    const idx_t num_tiles = (idx_t)1 << (log2cnt - P.log2b);
    const int log2ns = 63 - __clzll((unsigned long long)P.Ns);
    const idx_t ns_mask = (idx_t)P.Ns - 1;
    const idx_t tw_stride = (idx_t)P.tw_stride;
    const idx_t M32 = (idx_t)P.M;
    constexpr int PER = 8 / RAD;  // first-sub-pass butterflies per thread
    const int nbf = R / RAD;

    auto load = [&](idx_t n) {
        if (LOAD == LOAD_PLAIN) return P.in[n];
        double2 v = make_double2(0.0, 0.0);
        if (n < M32) {
            double d1, d2 = 0.0;
            if (LOAD == LOAD_ALPHA || LOAD == LOAD_ALPHA_TW) {
                alpha_signal(P.src, n, P.c1, P.c2, d1, d2);
            } else {
                d1 = __ldg(P.src + n);
            }
            const double2 c = __ldg(P.chirp + n);
            v = make_double2(d1 * c.x - d2 * c.y, d1 * c.y + d2 * c.x);
            if (LOAD == LOAD_ALPHA_TW)  // residue-1 input of a split transform: a[n] W_{2L}^n
                v = cmul(v, cmul(__ldg(P.rwhi + (n >> P.rsplit)), __ldg(P.rwlo + (n & (((idx_t)1 << P.rsplit) - 1)))));
        }
        return v;
    };
    // gather the 8 elements this thread feeds into the first sub-pass of tile `tl`
    auto gather = [&](idx_t tl, double2* dst) {
        const idx_t j_base = tl << P.log2b;
#pragma unroll
        for (int t = 0; t < PER; t++) {
            const int u = tid + t * T;
            const int i = u >> P.log2b, b = u & (B - 1);
#pragma unroll
            for (int q = 0; q < RAD; q++) dst[t * RAD + q] = load(j_base + b + ((idx_t)(i + q * nbf) << log2cnt));
        }
    };
    auto out_index = [&](idx_t j_base, int e, int& b, int& y) -> idx_t {
        if (P.Ns == 1) {  // out[j R + y]: R contiguous elements per batch member
            b = e >> P.log2r;
            y = e & (R - 1);
            return ((j_base + b) << P.log2r) + y;
        }
        // out[(j-k) R + k + y Ns]: B contiguous elements per y
        b = e & (B - 1);
        y = e >> P.log2b;
        const idx_t j = j_base + b;
        const idx_t k = j & ns_mask;
        return ((j - k) << P.log2r) + k + ((idx_t)y << log2ns);
    };

    for (idx_t tl = blockIdx.x; tl < num_tiles; tl += gridDim.x) {
        const idx_t j_base = tl << P.log2b;
        double2 v[8];
        if (active) gather(tl, v);

        // ---- first sub-pass (pass twiddle, radix RAD, no sub-twiddle) ----
        if (active) {
            auto tw = [&](idx_t m) {  // W_L^m (SGN-conjugated) from the two-level table
                const double2 a = __ldg(P.whi + (m >> P.split));
                const double2 bb = __ldg(P.wlo + (m & (((idx_t)1 << P.split) - 1)));
                double2 w = cmul(a, bb);
                if (SGN > 0) w.y = -w.y;
                return w;
            };
            const int b0 = tid & (B - 1);  // u = tid + t*T keeps the same batch member (T is a multiple of B)
            const idx_t k = (j_base + b0) & ns_mask;
            // element x = i + q*nbf carries W^{k x stride} = W^{k i stride} (W^{k nbf stride})^q:
            // one table lookup per butterfly plus one shared step instead of one lookup per element
            double2 wstep = make_double2(1.0, 0.0);
            if (P.Ns > 1) wstep = tw(k * (idx_t)nbf * tw_stride);
            // butterflies t = 0..PER-1 of a thread sit at i_t = i_0 + t (T >> log2b): their twiddles are
            // W^{k i_0 stride} (W^{k (T >> log2b) stride})^t — with PER > 2 one more lookup replaces PER - 1
            double2 wt_base = make_double2(1.0, 0.0), wt_step = make_double2(1.0, 0.0);
            if (PER > 2 && P.Ns > 1) {
                wt_base = tw(k * (idx_t)(tid >> P.log2b) * tw_stride);
                wt_step = tw(k * (idx_t)(T >> P.log2b) * tw_stride);
            }
#pragma unroll
            for (int t = 0; t < PER; t++) {
                const int u = tid + t * T;
                const int i = u >> P.log2b, b = u & (B - 1);
                if (P.Ns > 1) {
                    double2 w;
                    if (PER > 2) {
                        w = wt_base;
                        if (t + 1 < PER) wt_base = cmul(wt_base, wt_step);
                    } else {
                        w = tw(k * (idx_t)i * tw_stride);
                    }
#pragma unroll
                    for (int q = 0; q < RAD; q++) {
                        v[t * RAD + q] = cmul(v[t * RAD + q], w);
                        if (q + 1 < RAD) w = cmul(w, wstep);
                    }
                }
                if (RAD == 8) dft8<SGN>(v + RAD * t);
                if (RAD == 4) dft4<SGN>(v + RAD * t);
                if (RAD == 2) dft2<SGN>(v + RAD * t);
#pragma unroll
                for (int q = 0; q < RAD; q++) S[(i * RAD + q) * pitch + b] = v[t * RAD + q];
            }
        }
        double2 hv[8];
        if (STORE == STORE_MULH && active) {  // filter values for the store phase: issue early
#pragma unroll
            for (int t = 0; t < 8; t++) {
                int b, y;
                hv[t] = __ldg(P.H + out_index(j_base, tid + t * T, b, y));
            }
        }
        __syncthreads();
        // ---- remaining radix-8 sub-passes, in place through registers ----
        const int nb = R >> 3;
        constexpr int LASTF = (LAST == 8) ? 1 : LAST;
        for (int ns = RAD; ns * LASTF < R; ns <<= 3) {
            int i = 0, b = 0, kk = 0;
            if (active) {
                i = tid >> P.log2b;
                b = tid & (B - 1);
                kk = i & (ns - 1);
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = S[(i + q * nb) * pitch + b];
                double2 w1 = __ldg(P.wsub + kk * (4096 / (ns * 8)));
                if (SGN > 0) w1.y = -w1.y;
                double2 w = w1;
#pragma unroll
                for (int q = 1; q < 8; q++) {
                    v[q] = cmul(v[q], w);
                    if (q < 7) w = cmul(w, w1);
                }
                dft8<SGN>(v);
            }
            __syncthreads();
            if (active) {
#pragma unroll
                for (int q = 0; q < 8; q++) S[((i - kk) * 8 + kk + q * ns) * pitch + b] = v[q];
            }
            __syncthreads();
        }
        // ---- final radix-LAST sub-pass: in place (output index k + q Ns equals the input index) ----
        if (LAST != 8) {
            constexpr int PERL = 8 / LAST;
            const int nsl = R / LAST;
            if (active) {
#pragma unroll
                for (int t = 0; t < PERL; t++) {
                    const int u = tid + t * T;
                    const int i = u >> P.log2b, b = u & (B - 1);
                    double2* sp = S + i * pitch + b;
                    double2 x[LAST];
#pragma unroll
                    for (int q = 0; q < LAST; q++) x[q] = sp[q * nsl * pitch];
                    double2 w1 = __ldg(P.wsub + i * (4096 >> P.log2r));  // W_R^i
                    if (SGN > 0) w1.y = -w1.y;
                    double2 w = w1;
#pragma unroll
                    for (int q = 1; q < LAST; q++) {
                        x[q] = cmul(x[q], w);
                        if (q + 1 < LAST) w = cmul(w, w1);
                    }
                    if (LAST == 4) dft4<SGN>(x);
                    if (LAST == 2) dft2<SGN>(x);
#pragma unroll
                    for (int q = 0; q < LAST; q++) sp[q * nsl * pitch] = x[q];
                }
            }
            __syncthreads();
        }
        // ---- scatter to global ----
        if (active) {
#pragma unroll
            for (int t = 0; t < 8; t++) {
                int b, y;
                const idx_t gidx = out_index(j_base, tid + t * T, b, y);
                double2 val = S[y * pitch + b];
                if (STORE == STORE_PLAIN) {
                    P.out[gidx] = val;
                } else if (STORE == STORE_MULH) {
                    P.out[gidx] = cmul(val, hv[t]);
                } else if (gidx < M32) {
                    P.spec[gidx] = cmul(val, __ldg(P.chirp + gidx));
                }
            }
        }
        __syncthreads();  // S is rewritten by the next tile
    }
}

template <int SGN, int LOAD, int STORE, int LOG2TILE>
__global__ void __launch_bounds__((1 << LOG2TILE) / 8, LOG2TILE == kLog2TileSmall ? (STORE == STORE_MULH ? 4 : 7) : (STORE == STORE_MULH ? 2 : 3))
    fft_tile_kernel(PassParams P) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2* S = reinterpret_cast<double2*>(fft_smem);
    const int rem = P.log2r % 3;
    // Small tiles run 7 CTAs/SM on 72 registers: they keep the small radix in the FIRST sub-pass (one
    // code path less, no spills; measured 0.423 vs 0.429 ms at S = 1e6).  Large tiles (80 registers)
    // put it last: 2.80 vs 2.91 ms at S = 8e6.
    if (LOG2TILE == kLog2TileSmall) {  // (large tiles are only planned for L >= 2^12: every R >= 64)
        if (rem == 0) fft_tile_body<SGN, LOAD, STORE, 8, 8>(P, S);
        else if (rem == 2) fft_tile_body<SGN, LOAD, STORE, 4, 8>(P, S);
        else fft_tile_body<SGN, LOAD, STORE, 2, 8>(P, S);
    } else if (rem == 0) {
        fft_tile_body<SGN, LOAD, STORE, 8, 8>(P, S);
    } else if (rem == 2) {
        fft_tile_body<SGN, LOAD, STORE, 8, 4>(P, S);
    } else {
        fft_tile_body<SGN, LOAD, STORE, 8, 2>(P, S);
    }
}

struct FftIo {
    int load = LOAD_PLAIN;
    int store = STORE_PLAIN;
    const double2* H = nullptr;
    const double* src = nullptr;
    int c1 = 0, c2 = 0;
    double2* spec = nullptr;
    const double2* rwhi = nullptr;
    const double2* rwlo = nullptr;
    int rsplit = 0;
};

template <int SGN, int LOAD, int STORE, int LOG2TILE>
static int launch_tile_pass_t(const rn_spectrum_plan* p, const PassParams& P, cudaStream_t stream) {
    const int R = 1 << P.log2r, B = 1 << P.log2b;
    const int pitch = (B > 1) ? B + 1 : 1;
    const size_t smem = (size_t)R * pitch * sizeof(double2);
    const int64_t tiles = (P.L >> P.log2r) >> P.log2b;
    auto kern = fft_tile_kernel<SGN, LOAD, STORE, LOG2TILE>;
    (void)p;
    if (smem > 48 * 1024) RN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = std::min<int64_t>(tiles, (int64_t)1 << 30);  // one tile per CTA
    kern<<<(unsigned)grid, (1 << LOG2TILE) / 8, smem, stream>>>(P);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

template <int SGN, int LOAD, int STORE>
static int launch_tile_pass(const rn_spectrum_plan* p, const PassParams& P, cudaStream_t stream) {
    if (p->log2tile == kLog2TileLarge) return launch_tile_pass_t<SGN, LOAD, STORE, kLog2TileLarge>(p, P, stream);
    return launch_tile_pass_t<SGN, LOAD, STORE, kLog2TileSmall>(p, P, stream);
}

// Full length-L transform: `first` is read by the first pass (with io.load); passes ping-pong
// between the plan buffers (never writing the buffer being read); the last pass writes `last_out`
// if given (must not be a plan work buffer), else the next ping-pong buffer.  *result (optional)
// receives the buffer holding the output.
template <int SGN>
static int fft_run(const rn_spectrum_plan* p, const double2* first, double2* last_out, const FftIo& io,
                   cudaStream_t stream, double2** result = nullptr) {
    int64_t Ns = 1;
    const double2* src = first;
    for (int i = 0; i < p->num_passes; i++) {
        const bool is_first = (i == 0), is_last = (i == p->num_passes - 1);
        PassParams P;
        P.L = p->L;
        P.log2r = p->pass_log2r[i];
        int log2b = p->log2tile - P.log2r;
        const int log2cnt = p->log2l - P.log2r;
        if (log2b > log2cnt) log2b = log2cnt;
        if (log2b < 0) log2b = 0;
        P.log2b = log2b;
        P.Ns = Ns;
        P.tw_stride = p->L / (Ns << P.log2r);
        P.split = p->split;
        P.whi = p->d_whi;
        P.wlo = p->d_wlo;
        P.wsub = p->d_wsub;
        P.in = src;
        double2* dst = (is_last && last_out) ? last_out : ((src == p->d_buf0) ? p->d_buf1 : p->d_buf0);
        P.out = dst;
        P.H = io.H;
        P.src = io.src;
        P.c1 = io.c1;
        P.c2 = io.c2;
        P.M = p->M;
        P.chirp = p->d_chirp;
        P.spec = io.spec;
        P.rwhi = io.rwhi;
        P.rwlo = io.rwlo;
        P.rsplit = io.rsplit;
        const int load = is_first ? io.load : LOAD_PLAIN;
        const int store = is_last ? io.store : STORE_PLAIN;
        int rc = RN_ERR_UNSUPPORTED;
#define RN_PASS(LD, ST) \
    if (load == LD && store == ST) rc = launch_tile_pass<SGN, LD, ST>(p, P, stream);
        RN_PASS(LOAD_PLAIN, STORE_PLAIN)
        RN_PASS(LOAD_PLAIN, STORE_POST)
        RN_PASS(LOAD_PLAIN, STORE_MULH)
        RN_PASS(LOAD_ALPHA, STORE_PLAIN)
        RN_PASS(LOAD_ALPHA, STORE_MULH)
        RN_PASS(LOAD_SIGNAL, STORE_PLAIN)
        RN_PASS(LOAD_SIGNAL, STORE_MULH)
        RN_PASS(LOAD_ALPHA_TW, STORE_PLAIN)
        RN_PASS(LOAD_ALPHA_TW, STORE_MULH)
#undef RN_PASS
        if (rc != RN_OK) {
            if (rc == RN_ERR_UNSUPPORTED) set_error("unsupported FFT pass configuration");
            return rc;
        }
        Ns <<= P.log2r;
        src = dst;
        if (result) *result = dst;
    }
    return RN_OK;
}

// ---- Bluestein helper kernels ---------------------------------------------------------------

// chirp table c[n] = exp(-i*pi*n^2/M), n < M
__global__ void chirp_table_kernel(double2* __restrict__ out, int64_t M) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < M; n += (int64_t)gridDim.x * blockDim.x)
        out[n] = chirp(n, M);
}

// h[m] = exp(+i*pi*m^2/M) for |m| < M, stored circularly in a length-L array
__global__ void chirp_filter_kernel(double2* __restrict__ out, const double2* __restrict__ chirp_table, int64_t M,
                                    int64_t L) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < L; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t m = -1;
        if (idx < M) m = idx;
        else if (L - idx < M) m = L - idx;
        double2 v = make_double2(0.0, 0.0);
        if (m >= 0) {
            const double2 c = chirp_table[m];
            v = make_double2(c.x, -c.y);
        }
        out[idx] = v;
    }
}

// Energies sum_n s_n^2 of the seven signals of measure() (MODE 0) or of one real signal (MODE 1).
// Deterministic two-level reduction: per-block partials, then one block.
template <int MODE>
__global__ void __launch_bounds__(256) energy_partial_kernel(const double* __restrict__ src, int64_t n_begin,
                                                             int64_t M, double* __restrict__ partial) {
    double e[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t n = n_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < M;
         n += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == 0) {
            const double* a = src + n * 9;
            const double xx = a[9] - a[0], yy = a[13] - a[4], zz = a[17] - a[8];
            const double xy = a[10] - a[1], yz = a[14] - a[5], xz = a[11] - a[2];
            const double tr = xx + yy + zz, dxy = xx - yy, dyz = yy - zz, dzx = zz - xx;
            e[0] += tr * tr;
            e[1] += dxy * dxy;
            e[2] += dyz * dyz;
            e[3] += dzx * dzx;
            e[4] += xy * xy;
            e[5] += yz * yz;
            e[6] += xz * xz;
        } else {
            e[0] += src[n] * src[n];
        }
    }
    __shared__ double sm[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 7; q++) {
        double v = e[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sm[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) energy_final_kernel(const double* __restrict__ partial, int blocks,
                                                           double* __restrict__ energy) {
    // fixed-shape tree: thread t sums partials t, t+256, ...; then warp shuffles; then 8 warps
    __shared__ double sm[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 7; q++) {
        double v = 0;
        for (int b = threadIdx.x; b < blocks; b += 256) v += partial[(int64_t)b * 8 + q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sm[w][threadIdx.x];
        energy[threadIdx.x] = v;
    }
}

// The energies enter every bin of the orientational average as the same additive constant
// 45 (E_tr/9)/2 + 7 ((E_a + E_b + E_c)/2 + 3 (E_xy + E_yz + E_xz))/2; this is that constant for the
// frames the partials cover (sum it over shards, then add it to every bin).
__global__ void __launch_bounds__(256) energy_constant_kernel(const double* __restrict__ partial, int blocks,
                                                              double* __restrict__ out) {
    __shared__ double sm[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 7; q++) {
        double v = 0;
        for (int b = threadIdx.x; b < blocks; b += 256) v += partial[(int64_t)b * 8 + q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e[7];
        for (int q = 0; q < 7; q++) {
            double v = 0;
            for (int w = 0; w < 8; w++) v += sm[w][q];
            e[q] = v;
        }
        out[0] = 2.5 * e[0] + 1.75 * (e[1] + e[2] + e[3]) + 10.5 * (e[4] + e[5] + e[6]);
    }
}

struct SpectrumParams {
    double timestep;
    int laser;
    double laser_wavenumber;
    int bose_einstein;
    double kt;  // BOLTZMANN_CONSTANT * temperature
};

// wavenumbers: scipy.fftpack.fftfreq(M, dt)[k] * 33.35640951981521 * 1e3 with fftfreq = k * (1/(M*dt))
__device__ __forceinline__ double wavenumber_of(int64_t k, int64_t M, double dt) {
    const double val = 1.0 / ((double)M * dt);
    return ((double)k * val) * 33.35640951981521 * 1e3;
}

// Orientational average 45 a^2 + 7 g^2 (_raman.py:286-297) + corrections (_raman.py:13-69,303-307).
__global__ void __launch_bounds__(256) combine_kernel(const double2* __restrict__ spec, const double* __restrict__ energy,
                                                      int64_t M, int64_t L, int64_t points, SpectrumParams prm,
                                                      double* __restrict__ wn_out, double* __restrict__ int_out) {
    const double scale = 1.0 / (double)L;
    const double e_tr = energy[0], e_a = energy[1], e_b = energy[2], e_c = energy[3];
    const double e_xy = energy[4], e_yz = energy[5], e_xz = energy[6];
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < points; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = o + 1;  // bin 0 is dropped (_raman.py:299-301)
        double2 x[6];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double2 zk = spec[b * M + k], zm = spec[b * M + (M - k)];
            zk.x *= scale; zk.y *= scale; zm.x *= scale; zm.y *= scale;
            // z = x1 + i x2 with x1, x2 real signals: X1[k] = (Z[k] + conj Z[M-k])/2, X2[k] = (Z[k] - conj Z[M-k])/(2i)
            x[2 * b] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double dx = zk.x - zm.x, dy = zk.y + zm.y;
            x[2 * b + 1] = make_double2(0.5 * dy, -0.5 * dx);
        }
        const double2 xx = x[0], yy = x[1], zz = x[2], xy = x[3], yz = x[4], xz = x[5];
        auto power = [](double2 v, double e) { return (v.x * v.x + v.y * v.y + e) * 0.5; };
        const double s_tr = power(cadd(cadd(xx, yy), zz), e_tr);
        const double s_a = power(csub(xx, yy), e_a);
        const double s_b = power(csub(yy, zz), e_b);
        const double s_c = power(csub(zz, xx), e_c);
        const double s_xy = power(xy, e_xy), s_yz = power(yz, e_yz), s_xz = power(xz, e_xz);
        const double alpha2 = (1.0 / 9.0) * s_tr;
        const double gamma2 = (1.0 / 2.0) * s_a + (1.0 / 2.0) * s_b + (1.0 / 2.0) * s_c + 3.0 * s_xy + 3.0 * s_yz + 3.0 * s_xz;
        double inten = 45.0 * alpha2 + 7.0 * gamma2;
        const double wn = wavenumber_of(k, M, prm.timestep);
        if (prm.laser) {
            const double r = (wn - prm.laser_wavenumber) / 10000.0;
            const double r2 = r * r;
            inten *= (r2 * r2) / wn;
        }
        if (prm.bose_einstein) {
            const double en = wn * 29979245800.0 * 4.1357e-15;
            inten *= 1.0 / (1.0 - exp(-en / prm.kt));
        }
        wn_out[o] = wn;
        int_out[o] = inten;
    }
}

// calc_signal_spectrum for one real signal: I[k] = (|X[k]|^2 + E)/2, k = 0 .. ceil(M/2)-1
__global__ void __launch_bounds__(256) signal_combine_kernel(const double2* __restrict__ spec,
                                                             const double* __restrict__ energy, int64_t M, int64_t L,
                                                             int64_t points, double dt, double* __restrict__ wn_out,
                                                             double* __restrict__ int_out) {
    const double scale = 1.0 / (double)L;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < points; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 z = spec[k];
        const double re = z.x * scale, im = z.y * scale;
        int_out[k] = (re * re + im * im + energy[0]) * 0.5;
        wn_out[k] = wavenumber_of(k, M, dt);
    }
}

// One third of the orientational average, for the sharded (multi-GPU) measure: each part is one
// packed chirp-z transform, so parts can run on different ranks and their outputs just add up:
//   part 0: z = (xx-yy) + i (yy-zz)  ->  7 (S_a/2 + S_b/2 + S_c/2),  X_c = -(X_a + X_b)
//   part 1: z = trace + i xy         ->  45 S_tr/9 + 21 S_xy
//   part 2: z = yz + i xz            ->  21 (S_yz + S_xz)
__global__ void __launch_bounds__(256) part_combine_kernel(const double2* __restrict__ spec,
                                                           const double* __restrict__ energy, int part, int64_t M,
                                                           int64_t L, int64_t points, double* __restrict__ partial) {
    const double scale = 1.0 / (double)L;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < points; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = o + 1;
        double2 zk = spec[k], zm = spec[M - k];
        zk.x *= scale; zk.y *= scale; zm.x *= scale; zm.y *= scale;
        const double2 x1 = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        const double2 x2 = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
        auto power = [](double2 v, double e) { return (v.x * v.x + v.y * v.y + e) * 0.5; };
        double out;
        if (part == 0) {
            const double2 x3 = make_double2(-(x1.x + x2.x), -(x1.y + x2.y));
            out = 7.0 * ((1.0 / 2.0) * power(x1, energy[1]) + (1.0 / 2.0) * power(x2, energy[2]) +
                         (1.0 / 2.0) * power(x3, energy[3]));
        } else if (part == 1) {
            out = 45.0 * ((1.0 / 9.0) * power(x1, energy[0])) + 7.0 * (3.0 * power(x2, energy[4]));
        } else {
            out = 7.0 * (3.0 * power(x1, energy[5]) + 3.0 * power(x2, energy[6]));
        }
        partial[o] = out;
    }
}

// wavenumbers + optional corrections on summed partial intensities
__global__ void __launch_bounds__(256) finish_kernel(const double* __restrict__ partial_sum, int64_t M, int64_t points,
                                                     SpectrumParams prm, double* __restrict__ wn_out,
                                                     double* __restrict__ int_out) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < points; o += (int64_t)gridDim.x * blockDim.x) {
        double inten = partial_sum[o];
        const double wn = wavenumber_of(o + 1, M, prm.timestep);
        if (prm.laser) {
            const double r = (wn - prm.laser_wavenumber) / 10000.0;
            const double r2 = r * r;
            inten *= (r2 * r2) / wn;
        }
        if (prm.bose_einstein) {
            const double en = wn * 29979245800.0 * 4.1357e-15;
            inten *= 1.0 / (1.0 - exp(-en / prm.kt));
        }
        wn_out[o] = wn;
        int_out[o] = inten;
    }
}

static int grid_for(int64_t n, int sms) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16));
}

// chirp-z transform of one (packed) sequence: forward FFT with the Bluestein pre-multiply fused
// into the first pass' load, inverse FFT with the filter multiply fused into its first load and
// the chirp post-multiply fused into its last store.  Writes spec_out[0..M).
static int bluestein_transform(rn_spectrum_plan* p, const FftIo& in_io, double2* spec_out, cudaStream_t stream) {
    double2* fwd = nullptr;
    FftIo fio = in_io;
    fio.store = STORE_MULH;  // multiply by the filter spectrum while storing the forward transform
    fio.H = p->d_filter;
    int rc = fft_run<-1>(p, nullptr, nullptr, fio, stream, &fwd);
    if (rc != RN_OK) return rc;
    FftIo io;
    io.store = STORE_POST;
    io.spec = spec_out;
    return fft_run<+1>(p, fwd, nullptr, io, stream);
}

static void destroy_plan(rn_spectrum_plan* p) {
    if (!p) return;
    destroy_plan(p->half);
    cudaFree(p->d_hhalf[0]);
    cudaFree(p->d_hhalf[1]);
    cudaFree(p->d_buf0);
    cudaFree(p->d_buf1);
    cudaFree(p->d_filter);
    cudaFree(p->d_spec);
    if (p->owns_chirp) cudaFree(p->d_chirp);
    cudaFree(p->d_whi);
    cudaFree(p->d_wlo);
    cudaFree(p->d_wsub);
    cudaFree(p->d_partial);
    cudaFree(p->d_energy);
    delete p;
}

// exp(-2 pi i num/den) in 80-bit arithmetic, with exact octant symmetry handling left to cosl/sinl
static double2 unit_root(int64_t num, int64_t den) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    const long double ang = two_pi * (long double)num / (long double)den;
    return make_double2((double)cosl(ang), (double)-sinl(ang));
}

// pass structure + twiddle tables + work buffers of a length-2^log2l transform
static int setup_fft_core(rn_spectrum_plan* p, int log2l) {
    const int64_t L = (int64_t)1 << log2l;
    p->L = L;
    p->log2l = log2l;
    // sub-transforms of at most 256 points (>= 64-byte global chunks), as even as possible
    // (measured, tools/fft_pass_sweep.sh: two 512-point passes beat three for L = 2^17, 2^18 — 0.110 vs
    // 0.123 ms at S = 1e5; from 2^19 on the 16/32-byte gathers of larger radices lose)
    int max_log2r = (log2l == 17 || log2l == 18) ? 9 : kMaxLog2R;
    if (const char* env = getenv("RN_FFT_MAX_LOG2R")) max_log2r = atoi(env);  // tuning hook
    int large_min = kLargeTileMinLog2L;
    if (const char* env = getenv("RN_FFT_LARGE_TILE_MIN_LOG2L")) large_min = atoi(env);  // tuning hook
    p->log2tile = (log2l >= large_min && log2l >= 12) ? kLog2TileLarge : kLog2TileSmall;
    max_log2r = std::max(3, std::min(max_log2r, p->log2tile));
    p->num_passes = (log2l + max_log2r - 1) / max_log2r;
    for (int i = 0, rem = log2l; i < p->num_passes; i++) {
        const int left = p->num_passes - i;
        p->pass_log2r[i] = (rem + left - 1) / left;
        rem -= p->pass_log2r[i];
    }
    p->split = log2l / 2;
    const int64_t n_hi = L >> p->split, n_lo = (int64_t)1 << p->split;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (err == cudaSuccess) err = cudaMalloc(ptr, bytes);
    };
    alloc((void**)&p->d_buf0, sizeof(double2) * L);
    alloc((void**)&p->d_buf1, sizeof(double2) * L);
    alloc((void**)&p->d_whi, sizeof(double2) * n_hi);
    alloc((void**)&p->d_wlo, sizeof(double2) * n_lo);
    alloc((void**)&p->d_wsub, sizeof(double2) * 4096);
    if (err != cudaSuccess) {
        set_error("cudaMalloc failed while creating a length-2^%d transform: %s", log2l, cudaGetErrorString(err));
        cudaGetLastError();
        return RN_ERR_OUT_OF_MEMORY;
    }
    std::vector<double2> whi((size_t)n_hi), wlo((size_t)n_lo), wsub((size_t)4096);
    for (int64_t a = 0; a < n_hi; a++) whi[(size_t)a] = unit_root(a << p->split, L);
    for (int64_t b = 0; b < n_lo; b++) wlo[(size_t)b] = unit_root(b, L);
    for (int m = 0; m < 4096; m++) wsub[(size_t)m] = unit_root(m, 4096);
    cudaError_t e1 = cudaMemcpy(p->d_whi, whi.data(), sizeof(double2) * n_hi, cudaMemcpyHostToDevice);
    if (e1 == cudaSuccess) e1 = cudaMemcpy(p->d_wlo, wlo.data(), sizeof(double2) * n_lo, cudaMemcpyHostToDevice);
    if (e1 == cudaSuccess) e1 = cudaMemcpy(p->d_wsub, wsub.data(), sizeof(double2) * 4096, cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess) {
        set_error("twiddle table upload failed: %s", cudaGetErrorString(e1));
        return RN_ERR_CUDA;
    }
    return RN_OK;
}

// ---- transforms split over two ranks --------------------------------------------------------
// The chirp-z input a[n] = x[n] c[n] is zero for n >= M and L >= 2M-1, so the length-L transform
// splits by output residue r = k mod 2 into two length-L/2 transforms with no butterfly stage in
// front:  A[r + 2k'] = FFT_{L/2}( a[n] W_L^{n r} )[k'].  After the filter multiply (H[r + 2k']) and
// a length-L/2 inverse transform z_r, the chirp-z output is  y[m] = z_0[m] + W_L^{-m} z_1[m]
// (m < M <= L/2).  Two ranks each run one residue (half the work of a full transform), exchange
// nothing but reads of the partner's z_r over NVLink, and each finishes half of the bins.

__global__ void __launch_bounds__(256) decimate_filter_kernel(const double2* __restrict__ H, int64_t Lh,
                                                              double2* __restrict__ h0, double2* __restrict__ h1) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < Lh; k += (int64_t)gridDim.x * blockDim.x) {
        h0[k] = H[2 * k];
        h1[k] = H[2 * k + 1];
    }
}

// part_combine_kernel on the bins this residue's rank owns (residue 0: the lower half of the bins,
// residue 1: the upper half), from the two half-length inverse transforms
__global__ void __launch_bounds__(256) half_combine_kernel(const double2* __restrict__ z0, const double2* __restrict__ z1,
                                                           const double2* __restrict__ chirp_table,
                                                           const double2* __restrict__ whi,
                                                           const double2* __restrict__ wlo, int split,
                                                           const double* __restrict__ energy, int part, int residue,
                                                           int64_t M, int64_t L, int64_t points, int accumulate,
                                                           double* __restrict__ partial) {
    const double scale = 1.0 / (double)L;
    const int64_t cut = points / 2;  // bins o < cut belong to residue 0
    auto spec_at = [&](int64_t m) {
        double2 w = cmul(__ldg(whi + (m >> split)), __ldg(wlo + (m & (((int64_t)1 << split) - 1))));
        w.y = -w.y;  // W_L^{-m}
        const double2 y = cadd(z0[m], cmul(w, z1[m]));
        double2 v = cmul(y, __ldg(chirp_table + m));
        v.x *= scale;
        v.y *= scale;
        return v;
    };
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < points; o += (int64_t)gridDim.x * blockDim.x) {
        const bool mine = residue ? (o >= cut) : (o < cut);
        double out = 0.0;
        if (mine) {
            const int64_t k = o + 1;
            const double2 zk = spec_at(k), zm = spec_at(M - k);
            const double2 x1 = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 x2 = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
            auto power = [](double2 v, double e) { return (v.x * v.x + v.y * v.y + e) * 0.5; };
            if (part == 0) {
                const double2 x3 = make_double2(-(x1.x + x2.x), -(x1.y + x2.y));
                out = 7.0 * ((1.0 / 2.0) * power(x1, energy[1]) + (1.0 / 2.0) * power(x2, energy[2]) +
                             (1.0 / 2.0) * power(x3, energy[3]));
            } else if (part == 1) {
                out = 45.0 * ((1.0 / 9.0) * power(x1, energy[0])) + 7.0 * (3.0 * power(x2, energy[4]));
            } else {
                out = 7.0 * (3.0 * power(x1, energy[5]) + 3.0 * power(x2, energy[6]));
            }
        }
        if (accumulate) {
            if (mine) partial[o] += out;
        } else {
            partial[o] = out;
        }
    }
}

static int ensure_half_plan(rn_spectrum_plan* p) {
    if (p->half) return RN_OK;
    if (p->log2l < 4) {
        set_error("series too short for a split transform");
        return RN_ERR_UNSUPPORTED;
    }
    rn_spectrum_plan* h = new rn_spectrum_plan();
    h->device = p->device;
    h->sm_count = p->sm_count;
    h->S = p->S;
    h->M = p->M;
    h->d_chirp = p->d_chirp;
    h->owns_chirp = false;
    int rc = setup_fft_core(h, p->log2l - 1);
    if (rc == RN_OK) {
        cudaError_t err = cudaMalloc((void**)&p->d_hhalf[0], sizeof(double2) * h->L);
        if (err == cudaSuccess) err = cudaMalloc((void**)&p->d_hhalf[1], sizeof(double2) * h->L);
        if (err != cudaSuccess) {
            set_error("cudaMalloc failed for the decimated filter: %s", cudaGetErrorString(err));
            cudaGetLastError();
            rc = RN_ERR_OUT_OF_MEMORY;
        }
    }
    if (rc == RN_OK) {
        decimate_filter_kernel<<<grid_for(h->L, p->sm_count), 256>>>(p->d_filter, h->L, p->d_hhalf[0], p->d_hhalf[1]);
        RN_LAUNCHED();
        cudaError_t e2 = cudaStreamSynchronize(nullptr);
        if (e2 != cudaSuccess) {
            set_error("split-transform initialisation failed: %s", cudaGetErrorString(e2));
            rc = RN_ERR_CUDA;
        }
    }
    if (rc != RN_OK) {
        destroy_plan(h);
        cudaFree(p->d_hhalf[0]);
        cudaFree(p->d_hhalf[1]);
        p->d_hhalf[0] = p->d_hhalf[1] = nullptr;
        return rc;
    }
    p->half = h;
    return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" int64_t rn_spectrum_num_points(int64_t num_frames) {
    const int64_t M = num_frames - 1;
    if (M < 1) return 0;
    return (M + 1) / 2 - 1;
}

extern "C" int rn_spectrum_plan_create(int64_t num_frames, int device, rn_spectrum_plan** out) {
    RN_CHECK_ARG(out != nullptr, "out is null");
    *out = nullptr;
    RN_CHECK_ARG(num_frames >= 2, "a spectrum needs at least 2 frames (got %lld)", (long long)num_frames);
    int count = 0;
    RN_CUDA(cudaGetDeviceCount(&count));
    RN_CHECK_ARG(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("cudaSetDevice(%d) failed", device);
        return RN_ERR_CUDA;
    }
    cudaDeviceProp prop;
    RN_CUDA(cudaGetDeviceProperties(&prop, device));
    rn_spectrum_plan* p = new rn_spectrum_plan();
    p->device = device;
    p->sm_count = prop.multiProcessorCount;
    p->S = num_frames;
    p->M = num_frames - 1;
    int64_t L = 8;
    int log2l = 3;
    while (L < 2 * p->M - 1) {
        L <<= 1;
        log2l++;
    }
    RN_CHECK_ARG(log2l <= 30, "series too long for one spectrum plan (%lld frames)", (long long)num_frames);
    int core_rc = setup_fft_core(p, log2l);
    if (core_rc != RN_OK) {
        destroy_plan(p);
        return core_rc;
    }
    p->energy_blocks = (int)std::min<int64_t>((p->M + 255) / 256, (int64_t)p->sm_count * 4);
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (err == cudaSuccess) err = cudaMalloc(ptr, bytes);
    };
    alloc((void**)&p->d_filter, sizeof(double2) * L);
    alloc((void**)&p->d_spec, sizeof(double2) * 3 * p->M);
    alloc((void**)&p->d_chirp, sizeof(double2) * p->M);
    alloc((void**)&p->d_partial, sizeof(double) * 8 * p->energy_blocks);
    alloc((void**)&p->d_energy, sizeof(double) * 8);
    if (err != cudaSuccess) {
        set_error("cudaMalloc failed while creating a spectrum plan for %lld frames: %s", (long long)num_frames,
                  cudaGetErrorString(err));
        destroy_plan(p);
        cudaGetLastError();
        return RN_ERR_OUT_OF_MEMORY;
    }
    // chirp table, then the filter spectrum H = FFT_L(h) — computed once per plan
    chirp_table_kernel<<<grid_for(p->M, p->sm_count), 256>>>(p->d_chirp, p->M);
    RN_LAUNCHED();
    chirp_filter_kernel<<<grid_for(L, p->sm_count), 256>>>(p->d_buf0, p->d_chirp, p->M, L);
    RN_LAUNCHED();
    FftIo io;
    int rc = fft_run<-1>(p, p->d_buf0, p->d_filter, io, nullptr);
    if (rc == RN_OK) {
        cudaError_t e2 = cudaStreamSynchronize(nullptr);
        if (e2 != cudaSuccess) {
            set_error("spectrum plan initialisation failed: %s", cudaGetErrorString(e2));
            rc = RN_ERR_CUDA;
        }
    }
    if (rc != RN_OK) {
        destroy_plan(p);
        return rc;
    }
    *out = p;
    return RN_OK;
}

extern "C" int rn_spectrum_plan_destroy(rn_spectrum_plan* plan) {
    if (!plan) return RN_OK;
    DeviceGuard guard(plan->device);
    destroy_plan(plan);
    return RN_OK;
}

extern "C" int rn_md_spectrum(rn_spectrum_plan* plan, const double* d_alpha, double timestep_fs, int laser_correction,
                              double laser_wavelength_nm, int bose_einstein_correction, double temperature_K,
                              double* d_wavenumbers, double* d_intensities, void* stream) {
    RN_CHECK_ARG(plan != nullptr, "plan is null");
    RN_CHECK_ARG(d_alpha != nullptr, "d_alpha is null");
    RN_CHECK_ARG(timestep_fs > 0, "timestep must be positive");
    if (laser_correction) RN_CHECK_ARG(laser_wavelength_nm > 0, "invalid laser_wavelength");
    if (bose_einstein_correction) RN_CHECK_ARG(temperature_K > 0, "invalid temperature: %g <= 0", temperature_K);
    const int64_t points = rn_spectrum_num_points(plan->S);
    if (points == 0) return RN_OK;
    RN_CHECK_ARG(d_wavenumbers && d_intensities, "null output pointer");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t M = plan->M, L = plan->L;

    energy_partial_kernel<0><<<plan->energy_blocks, 256, 0, s>>>(d_alpha, 0, M, plan->d_partial);
    RN_LAUNCHED();
    energy_final_kernel<<<1, 256, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
    RN_LAUNCHED();
    // component pairs: (xx, yy), (zz, xy), (yz, xz)  — the upper triangle used at _raman.py:284-296
    const int pairs[3][2] = {{0, 4}, {8, 1}, {5, 2}};
    for (int b = 0; b < 3; b++) {
        FftIo io;
        io.load = LOAD_ALPHA;
        io.src = d_alpha;
        io.c1 = pairs[b][0];
        io.c2 = pairs[b][1];
        int rc = bluestein_transform(plan, io, plan->d_spec + (int64_t)b * M, s);
        if (rc != RN_OK) return rc;
    }
    SpectrumParams prm;
    prm.timestep = timestep_fs;
    prm.laser = laser_correction ? 1 : 0;
    prm.laser_wavenumber = laser_correction ? 10000000.0 / laser_wavelength_nm : 0.0;
    prm.bose_einstein = bose_einstein_correction ? 1 : 0;
    prm.kt = 8.617333262e-5 * temperature_K;  // constants.py:249
    combine_kernel<<<grid_for(points, plan->sm_count), 256, 0, s>>>(plan->d_spec, plan->d_energy, M, L, points, prm,
                                                                   d_wavenumbers, d_intensities);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

extern "C" int rn_md_spectrum_part(rn_spectrum_plan* plan, const double* d_alpha, int part, double* d_partial,
                                   void* stream) {
    RN_CHECK_ARG(plan != nullptr && d_alpha != nullptr, "null pointer");
    RN_CHECK_ARG(part >= 0 && part <= 2, "part must be 0, 1 or 2");
    const int64_t points = rn_spectrum_num_points(plan->S);
    if (points == 0) return RN_OK;
    RN_CHECK_ARG(d_partial != nullptr, "null output pointer");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t M = plan->M, L = plan->L;
    if (plan->energy_mode == 0) {
        energy_partial_kernel<0><<<plan->energy_blocks, 256, 0, s>>>(d_alpha, 0, M, plan->d_partial);
        RN_LAUNCHED();
        energy_final_kernel<<<1, 256, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
        RN_LAUNCHED();
    } else {
        RN_CUDA(cudaMemsetAsync(plan->d_energy, 0, sizeof(double) * 8, s));
    }
    FftIo io;
    io.load = LOAD_ALPHA;
    io.src = d_alpha;
    if (part == 0) io.c1 = -1;
    else if (part == 1) io.c1 = -2;
    else {
        io.c1 = 5;  // yz
        io.c2 = 2;  // xz
    }
    int rc = bluestein_transform(plan, io, plan->d_spec, s);
    if (rc != RN_OK) return rc;
    part_combine_kernel<<<grid_for(points, plan->sm_count), 256, 0, s>>>(plan->d_spec, plan->d_energy, part, M, L, points,
                                                                        d_partial);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// Length (complex elements) of the half transforms: rn_md_spectrum_half writes that many double2.
extern "C" int64_t rn_spectrum_half_length(const rn_spectrum_plan* plan) { return plan ? plan->L / 2 : 0; }

// One residue (0 or 1) of one packed transform (part 0..2) of measure(): writes the length-L/2
// inverse transform z_r to d_z_out (device memory other ranks can read, e.g. symmetric memory).
extern "C" int rn_md_spectrum_half(rn_spectrum_plan* plan, const double* d_alpha, int part, int residue,
                                   double* d_z_out, int skip_energy, void* stream) {
    RN_CHECK_ARG(plan != nullptr && d_alpha != nullptr && d_z_out != nullptr, "null pointer");
    RN_CHECK_ARG(part >= 0 && part <= 2, "part must be 0, 1 or 2");
    RN_CHECK_ARG(residue == 0 || residue == 1, "residue must be 0 or 1");
    if (rn_spectrum_num_points(plan->S) == 0) return RN_OK;
    DeviceGuard guard(plan->device);
    int rc = ensure_half_plan(plan);
    if (rc != RN_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rn_spectrum_plan* h = plan->half;
    if (plan->energy_mode != 0) {
        if (!skip_energy) RN_CUDA(cudaMemsetAsync(plan->d_energy, 0, sizeof(double) * 8, s));
    } else if (!skip_energy) {
        energy_partial_kernel<0><<<plan->energy_blocks, 256, 0, s>>>(d_alpha, 0, plan->M, plan->d_partial);
        RN_LAUNCHED();
        energy_final_kernel<<<1, 256, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
        RN_LAUNCHED();
    }
    // forward half transform: the chirp pre-multiply (and, residue 1, the W_L^n twiddle) in the first
    // pass' load, the decimated filter in the last pass' store
    double2* fwd = nullptr;
    FftIo fio;
    fio.load = residue ? LOAD_ALPHA_TW : LOAD_ALPHA;
    fio.src = d_alpha;
    if (part == 0) fio.c1 = -1;
    else if (part == 1) fio.c1 = -2;
    else {
        fio.c1 = 5;  // yz
        fio.c2 = 2;  // xz
    }
    fio.rwhi = plan->d_whi;
    fio.rwlo = plan->d_wlo;
    fio.rsplit = plan->split;
    fio.store = STORE_MULH;
    fio.H = plan->d_hhalf[residue];
    rc = fft_run<-1>(h, nullptr, nullptr, fio, s, &fwd);
    if (rc != RN_OK) return rc;
    FftIo io;
    rc = fft_run<+1>(h, fwd, reinterpret_cast<double2*>(d_z_out), io, s);
    if (rc != RN_OK) return rc;
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// Finishes the bins owned by `residue` (0: lower half, 1: upper half) of part `part` from the two
// residues' inverse transforms (one of them usually the partner rank's, read over NVLink); other bins
// are set to zero (accumulate == 0) or left alone.  The partial intensities of all parts and both
// residues add up to what rn_md_spectrum_part produces for the three parts.  Must run after
// rn_md_spectrum_half on the same plan (series energies).
extern "C" int rn_md_spectrum_half_combine(rn_spectrum_plan* plan, int part, int residue, const double* d_z_res0,
                                           const double* d_z_res1, double* d_partial, int accumulate, void* stream) {
    RN_CHECK_ARG(plan != nullptr, "plan is null");
    RN_CHECK_ARG(part >= 0 && part <= 2, "part must be 0, 1 or 2");
    RN_CHECK_ARG(residue == 0 || residue == 1, "residue must be 0 or 1");
    const int64_t points = rn_spectrum_num_points(plan->S);
    if (points == 0) return RN_OK;
    RN_CHECK_ARG(d_z_res0 && d_z_res1 && d_partial, "null pointer");
    DeviceGuard guard(plan->device);
    half_combine_kernel<<<grid_for(points, plan->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double2*>(d_z_res0), reinterpret_cast<const double2*>(d_z_res1), plan->d_chirp,
        plan->d_whi, plan->d_wlo, plan->split, plan->d_energy, part, residue, plan->M, plan->L, points, accumulate,
        d_partial);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// Multi-GPU measure: the series energies (one full pass over the series) shard over ranks.  With
// mode 1 rn_md_spectrum_part / rn_md_spectrum_half leave the energies out; rn_series_energy_constant
// writes the additive constant contributed by the difference signals n in [n_begin, n_end) (a subset of
// [0, S-1)) to d_out[0].  Summed over shards (it can ride as one extra element of the partial-intensity
// all-reduce) and added to every bin it restores exactly what mode 0 computes.
extern "C" int rn_spectrum_set_energy_mode(rn_spectrum_plan* plan, int mode) {
    RN_CHECK_ARG(plan != nullptr, "plan is null");
    RN_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    plan->energy_mode = mode;
    return RN_OK;
}

extern "C" int rn_series_energy_constant(rn_spectrum_plan* plan, const double* d_alpha, int64_t n_begin,
                                         int64_t n_end, double* d_out, void* stream) {
    RN_CHECK_ARG(plan != nullptr && d_alpha != nullptr && d_out != nullptr, "null pointer");
    RN_CHECK_ARG(n_begin >= 0 && n_begin <= n_end && n_end <= plan->M, "invalid range [%lld, %lld) of %lld",
                 (long long)n_begin, (long long)n_end, (long long)plan->M);
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    energy_partial_kernel<0><<<plan->energy_blocks, 256, 0, s>>>(d_alpha, n_begin, n_end, plan->d_partial);
    RN_LAUNCHED();
    energy_constant_kernel<<<1, 256, 0, s>>>(plan->d_partial, plan->energy_blocks, d_out);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

extern "C" int rn_md_spectrum_finish(int64_t num_frames, const double* d_partial_sum, double timestep_fs,
                                     int laser_correction, double laser_wavelength_nm, int bose_einstein_correction,
                                     double temperature_K, double* d_wavenumbers, double* d_intensities, void* stream) {
    RN_CHECK_ARG(timestep_fs > 0, "timestep must be positive");
    if (laser_correction) RN_CHECK_ARG(laser_wavelength_nm > 0, "invalid laser_wavelength");
    if (bose_einstein_correction) RN_CHECK_ARG(temperature_K > 0, "invalid temperature: %g <= 0", temperature_K);
    const int64_t points = rn_spectrum_num_points(num_frames);
    if (points == 0) return RN_OK;
    RN_CHECK_ARG(d_partial_sum && d_wavenumbers && d_intensities, "null pointer");
    SpectrumParams prm;
    prm.timestep = timestep_fs;
    prm.laser = laser_correction ? 1 : 0;
    prm.laser_wavenumber = laser_correction ? 10000000.0 / laser_wavelength_nm : 0.0;
    prm.bose_einstein = bose_einstein_correction ? 1 : 0;
    prm.kt = 8.617333262e-5 * temperature_K;
    finish_kernel<<<grid_for(points, 148), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_partial_sum, num_frames - 1, points, prm, d_wavenumbers, d_intensities);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

extern "C" int rn_signal_spectrum(rn_spectrum_plan* plan, const double* d_signal, double sampling_rate,
                                  double* d_wavenumbers, double* d_intensities, void* stream) {
    RN_CHECK_ARG(plan != nullptr, "plan is null");
    RN_CHECK_ARG(d_signal && d_wavenumbers && d_intensities, "null device pointer");
    RN_CHECK_ARG(sampling_rate > 0, "sampling_rate must be positive");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t M = plan->M, L = plan->L;
    const int64_t points = (M + 1) / 2;
    energy_partial_kernel<1><<<plan->energy_blocks, 256, 0, s>>>(d_signal, 0, M, plan->d_partial);
    RN_LAUNCHED();
    energy_final_kernel<<<1, 256, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
    RN_LAUNCHED();
    FftIo io;
    io.load = LOAD_SIGNAL;
    io.src = d_signal;
    int rc = bluestein_transform(plan, io, plan->d_spec, s);
    if (rc != RN_OK) return rc;
    signal_combine_kernel<<<grid_for(points, plan->sm_count), 256, 0, s>>>(plan->d_spec, plan->d_energy, M, L, points,
                                                                          sampling_rate, d_wavenumbers, d_intensities);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// Test hook (not in the public header): plain length-L transform of plan->L complex values,
// sign -1 forward / +1 inverse (unscaled).  d_out must not alias d_in.
extern "C" int rn_debug_fft(rn_spectrum_plan* plan, const double* d_in, double* d_out, int sign, void* stream) {
    RN_CHECK_ARG(plan && d_in && d_out, "null pointer");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    RN_CUDA(cudaMemcpyAsync(plan->d_buf1, d_in, sizeof(double2) * plan->L, cudaMemcpyDeviceToDevice, s));
    FftIo io;
    // the first pass reads d_buf1; intermediate passes ping-pong and never write the buffer being read
    if (sign < 0) return fft_run<-1>(plan, plan->d_buf1, reinterpret_cast<double2*>(d_out), io, s);
    return fft_run<+1>(plan, plan->d_buf1, reinterpret_cast<double2*>(d_out), io, s);
}
