// Spectrum smearing (convolve_spectrum, ramannoodle/spectrum/utils.py:12-73) and the
// host-buffer / pinned-memory entry points.
//
// out[l] = sum_i I_i f(wn_i - o_l), f = gaussian (1/w)(1/sqrt(2 pi)) exp(-x^2/(2 w^2)) or
// lorentzian (1/pi) (w/2) / (x^2 + (w/2)^2).  The reference loops over the K inputs in Python
// (O(K L) numpy work).  Here the (input chunk) x (output tile) grid is evaluated in
// parallel; each CTA first reduces the min/max of its input chunk and output tile and skips
// the pair when every Gaussian factor would underflow to exactly 0 in fp64 (|x| > 39 w:
// exp(-760.5) == 0), which is bit-identical to evaluating it.  Partial sums per input chunk
// are combined in a fixed order (deterministic).
#include <cfloat>

#include "rn_common.cuh"

namespace rn {

constexpr int kOutTile = 256;    // outputs per CTA (one per thread)
constexpr int kInChunk = 2048;   // inputs per CTA

__device__ __forceinline__ double block_reduce_minmax(double v, bool is_max, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, off);
        v = is_max ? fmax(v, o) : fmin(v, o);
    }
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = is_max ? fmax(r, sm[w]) : fmin(r, sm[w]);
    __syncthreads();
    return r;
}

template <int KIND>
__global__ void __launch_bounds__(kOutTile) smear_partial_kernel(const double* __restrict__ wn,
                                                                 const double* __restrict__ inten, int64_t K,
                                                                 const double* __restrict__ out_wn, int64_t L,
                                                                 double width, double* __restrict__ partial) {
    __shared__ double s_wn[kInChunk];
    __shared__ double s_in[kInChunk];
    __shared__ double s_red[kOutTile / 32];
    const int64_t l = (int64_t)blockIdx.x * kOutTile + threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * kInChunk;
    const int n_in = (int)min((int64_t)kInChunk, K - i0);
    double lo_in = DBL_MAX, hi_in = -DBL_MAX;
    bool finite = true;
    for (int i = threadIdx.x; i < n_in; i += kOutTile) {
        const double w = wn[i0 + i];
        s_wn[i] = w;
        s_in[i] = inten[i0 + i];
        lo_in = fmin(lo_in, w);
        hi_in = fmax(hi_in, w);
        finite = finite && (w - w == 0.0) && (s_in[i] - s_in[i] == 0.0);
    }
    const double o = (l < L) ? out_wn[l] : 0.0;
    double acc = 0.0;
    bool skip = false;
    if (KIND == 0) {
        // the pair can be skipped only if every factor underflows to +0 and nothing is NaN/Inf
        const double lo_o = block_reduce_minmax((l < L) ? o : DBL_MAX, false, s_red);
        const double hi_o = block_reduce_minmax((l < L) ? o : -DBL_MAX, true, s_red);
        lo_in = block_reduce_minmax(lo_in, false, s_red);
        hi_in = block_reduce_minmax(hi_in, true, s_red);
        const int all_finite = __syncthreads_and(finite && (o - o == 0.0));
        const double cut = 39.0 * width;
        skip = all_finite && (lo_in - hi_o > cut || lo_o - hi_in > cut);
    } else {
        __syncthreads();
    }
    if (!skip && l < L) {
        if (KIND == 0) {
            const double norm = (1 / width) * (1 / sqrt(2 * 3.141592653589793));
            const double denom = 2 * (width * width);
            for (int i = 0; i < n_in; i++) {
                const double dx = s_wn[i] - o;
                acc += (norm * exp(-(dx * dx) / denom)) * s_in[i];
            }
        } else {
            const double hw = 0.5 * width;
            const double hw2 = hw * hw;
            const double inv_pi = 1 / 3.141592653589793;
            for (int i = 0; i < n_in; i++) {
                const double dx = s_wn[i] - o;
                acc += (inv_pi * (hw / (dx * dx + hw2))) * s_in[i];
            }
        }
    }
    if (l < L) partial[(int64_t)blockIdx.y * L + l] = acc;
}

__global__ void smear_reduce_kernel(const double* __restrict__ partial, int64_t chunks, int64_t L,
                                    double* __restrict__ out) {
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < L; l += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int64_t c = 0; c < chunks; c++) v += partial[c * L + l];
        out[l] = v;
    }
}

// pipelined host-buffer evaluation state
struct HostPipe {
    double* d_pos[2] = {nullptr, nullptr};
    double* d_alpha[2] = {nullptr, nullptr};
    cudaStream_t stream[2] = {nullptr, nullptr};
    void release() {
        for (int i = 0; i < 2; i++) {
            if (d_pos[i]) cudaFree(d_pos[i]);
            if (d_alpha[i]) cudaFree(d_alpha[i]);
            if (stream[i]) cudaStreamDestroy(stream[i]);
            d_pos[i] = d_alpha[i] = nullptr;
            stream[i] = nullptr;
        }
    }
    // no destructor: at process exit the CUDA context may already be gone
};

}  // namespace rn

using namespace rn;

extern "C" size_t rn_convolve_workspace_size(int64_t num_in, int64_t num_out) {
    if (num_in <= 0 || num_out <= 0) return 0;
    const int64_t chunks = (num_in + kInChunk - 1) / kInChunk;
    return (size_t)chunks * (size_t)num_out * sizeof(double);
}

extern "C" int rn_convolve_spectrum(const double* d_wavenumbers, const double* d_intensities, int64_t num_in, int kind,
                                    double width, const double* d_out_wavenumbers, int64_t num_out,
                                    double* d_out_intensities, void* d_workspace, void* stream) {
    RN_CHECK_ARG(kind == 0 || kind == 1, "unsupported convolution type: %d", kind);
    RN_CHECK_ARG(width > 0, "invalid width: %g <= 0", width);
    RN_CHECK_ARG(num_in >= 0 && num_out >= 0, "negative size");
    if (num_out == 0) return RN_OK;
    RN_CHECK_ARG(d_out_wavenumbers && d_out_intensities, "null output pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (num_in == 0) {
        RN_CUDA(cudaMemsetAsync(d_out_intensities, 0, sizeof(double) * num_out, s));
        return RN_OK;
    }
    RN_CHECK_ARG(d_wavenumbers && d_intensities && d_workspace, "null device pointer");
    const int64_t chunks = (num_in + kInChunk - 1) / kInChunk;
    RN_CHECK_ARG(chunks <= 65535, "too many input points for one call (%lld)", (long long)num_in);
    dim3 grid((unsigned)((num_out + kOutTile - 1) / kOutTile), (unsigned)chunks);
    double* partial = static_cast<double*>(d_workspace);
    if (kind == 0)
        smear_partial_kernel<0><<<grid, kOutTile, 0, s>>>(d_wavenumbers, d_intensities, num_in, d_out_wavenumbers,
                                                          num_out, width, partial);
    else
        smear_partial_kernel<1><<<grid, kOutTile, 0, s>>>(d_wavenumbers, d_intensities, num_in, d_out_wavenumbers,
                                                          num_out, width, partial);
    RN_LAUNCHED();
    const int rgrid = (int)std::min<int64_t>((num_out + 255) / 256, 148 * 8);
    smear_reduce_kernel<<<rgrid, 256, 0, s>>>(partial, chunks, num_out, d_out_intensities);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

extern "C" int rn_host_register(void* h_ptr, size_t bytes) {
    RN_CHECK_ARG(h_ptr != nullptr && bytes > 0, "invalid host buffer");
    RN_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
    return RN_OK;
}

extern "C" int rn_host_unregister(void* h_ptr) {
    RN_CHECK_ARG(h_ptr != nullptr, "invalid host buffer");
    RN_CUDA(cudaHostUnregister(h_ptr));
    return RN_OK;
}

extern "C" int rn_calc_polarizabilities_multi(const rn_model* model, const double* d_positions, int64_t num_frames,
                                              double* const* d_alpha_outputs, int num_outputs, void* stream);

// Staging buffers and streams of the host-buffer entries, cached per thread (allocation / free
// would serialise the device).
static int acquire_pipe(int device, int64_t pos_doubles, int64_t alpha_doubles, HostPipe** out) {
    static thread_local HostPipe pipe;
    static thread_local int pipe_device = -1;
    static thread_local int64_t pipe_pos_doubles = 0, pipe_alpha_doubles = 0;
    if (pipe_device != device) {
        pipe.release();
        pipe_device = device;
        pipe_pos_doubles = pipe_alpha_doubles = 0;
        for (int i = 0; i < 2; i++) RN_CUDA(cudaStreamCreateWithFlags(&pipe.stream[i], cudaStreamNonBlocking));
    }
    if (pipe_pos_doubles < pos_doubles) {
        for (int i = 0; i < 2; i++) {
            if (pipe.d_pos[i]) RN_CUDA(cudaFree(pipe.d_pos[i]));
            pipe.d_pos[i] = nullptr;
            RN_CUDA(cudaMalloc((void**)&pipe.d_pos[i], sizeof(double) * pos_doubles));
        }
        pipe_pos_doubles = pos_doubles;
    }
    if (pipe_alpha_doubles < alpha_doubles) {
        for (int i = 0; i < 2; i++) {
            if (pipe.d_alpha[i]) RN_CUDA(cudaFree(pipe.d_alpha[i]));
            pipe.d_alpha[i] = nullptr;
            RN_CUDA(cudaMalloc((void**)&pipe.d_alpha[i], sizeof(double) * alpha_doubles));
        }
        pipe_alpha_doubles = alpha_doubles;
    }
    *out = &pipe;
    return RN_OK;
}

// Shared implementation of the host-buffer entries: outputs[0] is the local series (or null when
// only h_alpha is wanted), further outputs are peer GPUs' buffers (fused all-gather).
static int host_pipeline(const rn_model* model, const double* h_positions, int64_t num_frames, double* h_alpha,
                         double* const* d_outputs, int num_outputs, int64_t chunk_frames) {
    RN_CHECK_ARG(model != nullptr, "model is null");
    RN_CHECK_ARG(num_frames >= 0, "num_frames must be non-negative");
    if (num_frames == 0) return RN_OK;
    RN_CHECK_ARG(h_positions != nullptr, "null host pointer");
    double* d_alpha = (num_outputs > 0) ? d_outputs[0] : nullptr;
    RN_CHECK_ARG(h_alpha || d_alpha, "no output buffer given");
    RN_CHECK_ARG(num_outputs <= 8, "at most 8 output pointers");
    DeviceGuard guard(model->device);
    if (!guard.ok) {
        set_error("cudaSetDevice(%d) failed", model->device);
        return RN_ERR_CUDA;
    }
    const int64_t K = model->dim;
    // ~256 MiB chunks: measured on B200 (tools/e2e_probe2.py) 64 MiB chunks show sporadic multi-100-ms
    // stalls, 256 MiB chunks run at a steady ~48 GB/s of the 55 GB/s pinned-copy rate
    if (chunk_frames <= 0) chunk_frames = std::max<int64_t>(8, (int64_t)(256ll << 20) / (K * 8));
    chunk_frames = std::min(chunk_frames, num_frames);
    chunk_frames = (chunk_frames + 7) / 8 * 8;
    HostPipe* pipe_ptr = nullptr;
    int prc = acquire_pipe(model->device, chunk_frames * K, d_alpha ? 0 : chunk_frames * 9, &pipe_ptr);
    if (prc != RN_OK) return prc;
    HostPipe& pipe = *pipe_ptr;
    int slot = 0;
    for (int64_t f0 = 0; f0 < num_frames; f0 += chunk_frames, slot ^= 1) {
        const int64_t n = std::min(chunk_frames, num_frames - f0);
        cudaStream_t s = pipe.stream[slot];
        // stream order serialises reuse of this slot's buffers; the other slot overlaps
        RN_CUDA(cudaMemcpyAsync(pipe.d_pos[slot], h_positions + f0 * K, sizeof(double) * n * K, cudaMemcpyHostToDevice, s));
        double* outs[8];
        int count = 1;
        outs[0] = d_alpha ? d_alpha + f0 * 9 : pipe.d_alpha[slot];
        for (int p = 1; p < num_outputs; p++) outs[count++] = d_outputs[p] + f0 * 9;
        int rc = rn_calc_polarizabilities_multi(model, pipe.d_pos[slot], n, outs, count, s);
        if (rc != RN_OK) return rc;
        if (h_alpha) RN_CUDA(cudaMemcpyAsync(h_alpha + f0 * 9, outs[0], sizeof(double) * n * 9, cudaMemcpyDeviceToHost, s));
    }
    RN_CUDA(cudaStreamSynchronize(pipe.stream[0]));
    RN_CUDA(cudaStreamSynchronize(pipe.stream[1]));
    return RN_OK;
}

extern "C" int rn_calc_polarizabilities_host(const rn_model* model, const double* h_positions, int64_t num_frames,
                                             double* h_alpha, double* d_alpha, int64_t chunk_frames) {
    double* outs[1] = {d_alpha};
    return host_pipeline(model, h_positions, num_frames, h_alpha, outs, d_alpha ? 1 : 0, chunk_frames);
}

extern "C" int rn_calc_polarizabilities_sweep(const rn_model* const* models, int num_models, const double* d_positions,
                                              int64_t num_frames, double* const* d_alpha_outputs, void* stream);

// Mask sweep from a host trajectory: every chunk crosses PCIe once and is evaluated by all models.
extern "C" int rn_calc_polarizabilities_host_sweep(const rn_model* const* models, int num_models,
                                                   const double* h_positions, int64_t num_frames,
                                                   double* const* d_alpha_outputs, int64_t chunk_frames) {
    RN_CHECK_ARG(models != nullptr && num_models >= 1 && num_models <= 64, "1..64 models are required");
    RN_CHECK_ARG(d_alpha_outputs != nullptr, "output pointers are required");
    for (int g = 0; g < num_models; g++) {
        RN_CHECK_ARG(models[g] != nullptr && d_alpha_outputs[g] != nullptr, "null model or output pointer");
        RN_CHECK_ARG(models[g]->device == models[0]->device && models[g]->dim == models[0]->dim,
                     "models of a sweep must live on one device and describe the same structure");
    }
    RN_CHECK_ARG(num_frames >= 0, "num_frames must be non-negative");
    if (num_frames == 0) return RN_OK;
    RN_CHECK_ARG(h_positions != nullptr, "null host pointer");
    const rn_model* model = models[0];
    DeviceGuard guard(model->device);
    if (!guard.ok) {
        set_error("cudaSetDevice(%d) failed", model->device);
        return RN_ERR_CUDA;
    }
    const int64_t K = model->dim;
    if (chunk_frames <= 0) chunk_frames = std::max<int64_t>(8, (int64_t)(256ll << 20) / (K * 8));
    chunk_frames = std::min(chunk_frames, num_frames);
    chunk_frames = (chunk_frames + 7) / 8 * 8;
    HostPipe* pipe = nullptr;
    int rc = acquire_pipe(model->device, chunk_frames * K, 0, &pipe);
    if (rc != RN_OK) return rc;
    int slot = 0;
    for (int64_t f0 = 0; f0 < num_frames; f0 += chunk_frames, slot ^= 1) {
        const int64_t n = std::min(chunk_frames, num_frames - f0);
        cudaStream_t s = pipe->stream[slot];
        RN_CUDA(cudaMemcpyAsync(pipe->d_pos[slot], h_positions + f0 * K, sizeof(double) * n * K, cudaMemcpyHostToDevice, s));
        double* outs[64];
        for (int g = 0; g < num_models; g++) outs[g] = d_alpha_outputs[g] + f0 * 9;
        rc = rn_calc_polarizabilities_sweep(models, num_models, pipe->d_pos[slot], n, outs, s);
        if (rc != RN_OK) return rc;
    }
    RN_CUDA(cudaStreamSynchronize(pipe->stream[0]));
    RN_CUDA(cudaStreamSynchronize(pipe->stream[1]));
    return RN_OK;
}

extern "C" int rn_calc_polarizabilities_host_multi(const rn_model* model, const double* h_positions, int64_t num_frames,
                                                   double* const* d_alpha_outputs, int num_outputs,
                                                   int64_t chunk_frames) {
    RN_CHECK_ARG(d_alpha_outputs != nullptr && num_outputs >= 1, "output pointers are required");
    for (int i = 0; i < num_outputs; i++) RN_CHECK_ARG(d_alpha_outputs[i] != nullptr, "null output pointer");
    return host_pipeline(model, h_positions, num_frames, nullptr, d_alpha_outputs, num_outputs, chunk_frames);
}
