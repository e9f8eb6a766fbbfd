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

#include "rn_common.cuh"

struct rn_spectrum_plan {
    int device = 0;
    int sm_count = 0;
    int64_t S = 0, M = 0, L = 0;
    double2* d_buf0 = nullptr;    // L
    double2* d_buf1 = nullptr;    // L
    double2* d_filter = nullptr;  // L   FFT of the chirp filter
    double2* d_spec = nullptr;    // 3*M  chirp-z outputs (unscaled by 1/L)
    double* d_partial = nullptr;  // energy_blocks * 8
    double* d_energy = nullptr;   // 8
    int energy_blocks = 0;
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

// One Stockham radix-R pass over a length-L sequence; Ns = product of the radices already done.
// SGN = -1 forward, +1 inverse (unscaled).  MULH multiplies the input by H (pointwise) first.
template <int R, int SGN, bool MULH>
__global__ void __launch_bounds__(256) fft_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                       const double2* __restrict__ H, int64_t L, int64_t Ns) {
    const int64_t count = L / R;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < count;
         j += (int64_t)gridDim.x * blockDim.x) {
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            v[r] = in[j + r * count];
            if (MULH) v[r] = cmul(v[r], H[j + r * count]);
        }
        const int64_t k = j & (Ns - 1);
        if (Ns > 1) {
            double s, c;
            sincospi(2.0 * (double)k / (double)(Ns * R), &s, &c);
            const double2 w1 = make_double2(c, SGN * s);
            double2 w = w1;
#pragma unroll
            for (int r = 1; r < R; r++) {
                v[r] = cmul(v[r], w);
                if (r + 1 < R) w = cmul(w, w1);
            }
        }
        if (R == 8) dft8<SGN>(v);
        if (R == 4) dft4<SGN>(v);
        if (R == 2) dft2<SGN>(v);
        const int64_t j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) out[j0 + r * Ns] = v[r];
    }
}

template <int SGN, bool MULH>
static int launch_pass(int R, const double2* in, double2* out, const double2* H, int64_t L, int64_t Ns, int sms,
                       cudaStream_t stream) {
    const int64_t count = L / R;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, (int64_t)sms * 16));
    if (R == 8) fft_pass_kernel<8, SGN, MULH><<<grid, 256, 0, stream>>>(in, out, H, L, Ns);
    if (R == 4) fft_pass_kernel<4, SGN, MULH><<<grid, 256, 0, stream>>>(in, out, H, L, Ns);
    if (R == 2) fft_pass_kernel<2, SGN, MULH><<<grid, 256, 0, stream>>>(in, out, H, L, Ns);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// Full length-L FFT, ping-ponging between a and b; *result receives the buffer holding the output.
template <int SGN>
static int fft_pow2(double2* a, double2* b, const double2* H, int64_t L, int sms, cudaStream_t stream,
                    double2** result) {
    int log2l = 0;
    while (((int64_t)1 << log2l) < L) log2l++;
    int64_t Ns = 1;
    double2 *src = a, *dst = b;
    bool first = true;
    int rem = log2l;
    while (rem > 0) {
        int R = 8;
        if (rem % 3 == 1) R = 2;       // take the odd factor first (twiddle-free while Ns == 1)
        else if (rem % 3 == 2) R = 4;
        int rc;
        if (first && H != nullptr) rc = launch_pass<SGN, true>(R, src, dst, H, L, Ns, sms, stream);
        else rc = launch_pass<SGN, false>(R, src, dst, nullptr, L, Ns, sms, stream);
        if (rc != RN_OK) return rc;
        first = false;
        Ns *= R;
        rem -= (R == 8) ? 3 : (R == 4 ? 2 : 1);
        std::swap(src, dst);
    }
    *result = src;
    return RN_OK;
}

// ---- Bluestein pre/post kernels -----------------------------------------------------------

// h[m] = exp(+i*pi*m^2/M) for |m| < M, stored circularly in a length-L array
__global__ void chirp_filter_kernel(double2* __restrict__ out, int64_t M, int64_t L) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < L; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t m = -1;
        if (idx < M) m = idx;
        else if (L - idx < M) m = L - idx;
        double2 v = make_double2(0.0, 0.0);
        if (m >= 0) {
            const double2 c = chirp(m, M);
            v = make_double2(c.x, -c.y);
        }
        out[idx] = v;
    }
}

// a[n] = (d1[n] + i d2[n]) * chirp(n) for n < M, 0 for M <= n < L, where d = diff of the
// polarizability series (np.diff, _raman.py:282) for tensor components (c1, c2).
// MODE 0: alpha series (stride 9, differences); MODE 1: a plain real signal (no diff, imag = 0).
template <int MODE>
__global__ void bluestein_prep_kernel(const double* __restrict__ src, int c1, int c2, double2* __restrict__ out,
                                      int64_t M, int64_t L) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < L; n += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (n < M) {
            double d1, d2;
            if (MODE == 0) {
                d1 = src[(n + 1) * 9 + c1] - src[n * 9 + c1];
                d2 = src[(n + 1) * 9 + c2] - src[n * 9 + c2];
            } else {
                d1 = src[n];
                d2 = 0.0;
            }
            const double2 c = chirp(n, M);
            v = make_double2(d1 * c.x - d2 * c.y, d1 * c.y + d2 * c.x);
        }
        out[n] = v;
    }
}

// spec[k] = chirp(k) * y[k], k < M   (the 1/L of the inverse FFT is applied by the consumer)
__global__ void bluestein_post_kernel(const double2* __restrict__ y, double2* __restrict__ spec, int64_t M) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x)
        spec[k] = cmul(y[k], chirp(k, M));
}

// Energies sum_n s_n^2 of the seven signals of measure() (MODE 0) or of one real signal (MODE 1).
// Deterministic two-level reduction: per-block partials, then one block.
template <int MODE>
__global__ void __launch_bounds__(256) energy_partial_kernel(const double* __restrict__ src, int64_t M,
                                                             double* __restrict__ partial) {
    double e[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < M; n += (int64_t)gridDim.x * blockDim.x) {
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

__global__ void energy_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ energy) {
    if (threadIdx.x < 7) {
        double v = 0;
        for (int b = 0; b < blocks; b++) v += partial[(int64_t)b * 8 + threadIdx.x];
        energy[threadIdx.x] = v;
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

static int grid_for(int64_t n, int sms) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16));
}

// chirp-z transform of the sequence already prepared in plan->d_buf0; writes spec_out[0..M)
static int bluestein_transform(rn_spectrum_plan* p, double2* spec_out, cudaStream_t stream) {
    double2* fwd = nullptr;
    int rc = fft_pow2<-1>(p->d_buf0, p->d_buf1, nullptr, p->L, p->sm_count, stream, &fwd);
    if (rc != RN_OK) return rc;
    double2* other = (fwd == p->d_buf0) ? p->d_buf1 : p->d_buf0;
    double2* inv = nullptr;
    rc = fft_pow2<+1>(fwd, other, p->d_filter, p->L, p->sm_count, stream, &inv);
    if (rc != RN_OK) return rc;
    bluestein_post_kernel<<<grid_for(p->M, p->sm_count), 256, 0, stream>>>(inv, spec_out, p->M);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

static void destroy_plan(rn_spectrum_plan* p) {
    if (!p) return;
    cudaFree(p->d_buf0);
    cudaFree(p->d_buf1);
    cudaFree(p->d_filter);
    cudaFree(p->d_spec);
    cudaFree(p->d_partial);
    cudaFree(p->d_energy);
    delete p;
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
    while (L < 2 * p->M - 1) L <<= 1;
    p->L = L;
    p->energy_blocks = (int)std::min<int64_t>((p->M + 255) / 256, (int64_t)p->sm_count * 4);
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (err == cudaSuccess) err = cudaMalloc(ptr, bytes);
    };
    alloc((void**)&p->d_buf0, sizeof(double2) * L);
    alloc((void**)&p->d_buf1, sizeof(double2) * L);
    alloc((void**)&p->d_filter, sizeof(double2) * L);
    alloc((void**)&p->d_spec, sizeof(double2) * 3 * p->M);
    alloc((void**)&p->d_partial, sizeof(double) * 8 * p->energy_blocks);
    alloc((void**)&p->d_energy, sizeof(double) * 8);
    if (err != cudaSuccess) {
        set_error("cudaMalloc failed while creating a spectrum plan for %lld frames: %s", (long long)num_frames,
                  cudaGetErrorString(err));
        destroy_plan(p);
        cudaGetLastError();
        return RN_ERR_OUT_OF_MEMORY;
    }
    // filter spectrum H = FFT_L(h), computed once per plan
    chirp_filter_kernel<<<grid_for(L, p->sm_count), 256>>>(p->d_buf0, p->M, L);
    RN_LAUNCHED();
    double2* res = nullptr;
    int rc = fft_pow2<-1>(p->d_buf0, p->d_buf1, nullptr, L, p->sm_count, nullptr, &res);
    if (rc == RN_OK) {
        cudaError_t e2 = cudaMemcpyAsync(p->d_filter, res, sizeof(double2) * L, cudaMemcpyDeviceToDevice, nullptr);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(nullptr);
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

    energy_partial_kernel<0><<<plan->energy_blocks, 256, 0, s>>>(d_alpha, M, plan->d_partial);
    RN_LAUNCHED();
    energy_final_kernel<<<1, 32, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
    RN_LAUNCHED();
    // component pairs: (xx, yy), (zz, xy), (yz, xz)  — the upper triangle used at _raman.py:284-296
    const int pairs[3][2] = {{0, 4}, {8, 1}, {5, 2}};
    for (int b = 0; b < 3; b++) {
        bluestein_prep_kernel<0><<<grid_for(L, plan->sm_count), 256, 0, s>>>(d_alpha, pairs[b][0], pairs[b][1],
                                                                            plan->d_buf0, M, L);
        RN_LAUNCHED();
        int rc = bluestein_transform(plan, plan->d_spec + (int64_t)b * M, s);
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

extern "C" int rn_signal_spectrum(rn_spectrum_plan* plan, const double* d_signal, double sampling_rate,
                                  double* d_wavenumbers, double* d_intensities, void* stream) {
    RN_CHECK_ARG(plan != nullptr, "plan is null");
    RN_CHECK_ARG(d_signal && d_wavenumbers && d_intensities, "null device pointer");
    RN_CHECK_ARG(sampling_rate > 0, "sampling_rate must be positive");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t M = plan->M, L = plan->L;
    const int64_t points = (M + 1) / 2;
    energy_partial_kernel<1><<<plan->energy_blocks, 256, 0, s>>>(d_signal, M, plan->d_partial);
    RN_LAUNCHED();
    energy_final_kernel<<<1, 32, 0, s>>>(plan->d_partial, plan->energy_blocks, plan->d_energy);
    RN_LAUNCHED();
    bluestein_prep_kernel<1><<<grid_for(L, plan->sm_count), 256, 0, s>>>(d_signal, 0, 0, plan->d_buf0, M, L);
    RN_LAUNCHED();
    int rc = bluestein_transform(plan, plan->d_spec, s);
    if (rc != RN_OK) return rc;
    signal_combine_kernel<<<grid_for(points, plan->sm_count), 256, 0, s>>>(plan->d_spec, plan->d_energy, M, L, points,
                                                                          sampling_rate, d_wavenumbers, d_intensities);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}
