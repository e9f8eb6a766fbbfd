// Polarizability evaluation kernels (sm_100a).
//
//   alpha_s = alpha_ref + sum_j (1-mask_j) B_j( v_j . vec(wrap(p_s - p_ref) L) )
//   (ramannoodle/pmodel/_interpolation.py:191-252, structure/utils.py:110-135,
//    structure/_reference.py:268-285)
//
// Three kernels:
//   affine_tma_kernel      linear DOFs collapsed to alpha0 + D.G.  Frames are streamed with
//                          TMA bulk copies (cp.async.bulk) into a 4-stage shared-memory ring;
//                          8 consumer warps each own a K-slice whose G / reference fragments
//                          stay in registers; the contraction runs on the FP64 tensor pipe
//                          (DMMA m8n8k4) for 8 of the 9 tensor components and as DFMA for
//                          the 9th.  HBM-bound: 24N+72 algorithmic bytes per frame.
//   affine_generic_kernel  same math, any size/alignment, no staging (fallback + cross-check).
//   dense_kernel           general splines: cp.async-pipelined DMMA GEMM (frames x 3N)(3N x J)
//                          with the minimum-image wrap applied while loading A fragments, and
//                          the piecewise-polynomial spline evaluation + 3x3 reduction fused as
//                          the epilogue (amplitudes never touch HBM).  FP64-pipe-bound:
//                          2*3N*J flops per frame.
#include "rn_device.cuh"

namespace rn {

// ------------------------------------------------------------------------------------
// generic affine kernel: one warp per group of 8 frames, A fragments straight from global
// ------------------------------------------------------------------------------------
template <bool WRAP>
__global__ void __launch_bounds__(256) affine_generic_kernel(const double* __restrict__ in,
                                                             const double* __restrict__ ref,
                                                             const double* __restrict__ G, int64_t frame_begin,
                                                             int64_t frame_end, int K, Alpha0 a0,
                                                             double* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t groups = (frame_end - frame_begin + 7) / 8;
    for (int64_t grp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); grp < groups; grp += warps) {
        const int64_t frame = frame_begin + grp * 8 + g;
        const bool valid = frame < frame_end;
        const double* row = in + (valid ? frame : frame_begin) * (int64_t)K;
        double c0 = 0, c1 = 0, d0 = 0, d1 = 0;
        for (int e0 = 0; e0 < K; e0 += 4) {
            const int e = e0 + t;
            double a = 0, b = 0, b2 = 0;
            if (e < K) {
                if (valid) {
                    a = __ldg(row + e);
                    if (WRAP) a = wrap_disp(a, __ldg(ref + e));
                }
                b = __ldg(G + (int64_t)e * 9 + g);
                if (g == 0) b2 = __ldg(G + (int64_t)e * 9 + 8);
            }
            dmma884(c0, c1, a, b);
            dmma884(d0, d1, a, b2);
        }
        if (valid) {
            alpha[frame * 9 + 2 * t] = c0 + a0.v[2 * t];
            alpha[frame * 9 + 2 * t + 1] = c1 + a0.v[2 * t + 1];
            if (t == 0) alpha[frame * 9 + 8] = d0 + a0.v[8];
        }
    }
}

// ------------------------------------------------------------------------------------
// TMA affine kernel
// ------------------------------------------------------------------------------------
constexpr int kRedStride = 12;  // doubles per (warp, frame) slot in the cross-warp reduction buffer

// Stage layout: 8*MT frame rows, each one bulk copy of K doubles, at a stride == 8 (mod 16)
// doubles so that the LDS.128 fragment loads of two consecutive rows hit disjoint banks.
// Requires K even (frame rows are then multiples of 16 bytes, as cp.async.bulk needs).
struct AffineSmemLayout {
    int row_stride;  // doubles
    int stage_doubles;
    size_t bytes;
};

static AffineSmemLayout affine_layout(int K, int kp, int mt, int stages) {
    AffineSmemLayout L;
    L.row_stride = K + (8 - K % 16 + 16) % 16;
    const int slack = kAffineWarps * 8 * kp;  // fragment loads may run past the last row (values are masked)
    L.stage_doubles = 8 * mt * L.row_stride;
    L.bytes = (size_t)(stages * L.stage_doubles + slack) * 8 + (size_t)2 * kAffineWarps * 8 * mt * kRedStride * 8 +
              (size_t)2 * slack * 8 + (size_t)2 * 8 * mt * 9 * 8 + (size_t)stages * 8 + 64;
    return L;
}

// 8 warps per CTA, two CTAs per SM when registers allow (while one CTA sits in its reduction /
// barrier phase the other one computes).  A tile is 8*MT consecutive frames; STAGES bulk-copy
// tiles are kept in flight per CTA.  Every warp contracts its own K-slice (8*KP elements) of all
// rows of the tile: the DMMA B fragments (8 tensor components) stay in registers, the 9th
// component's column of G and the reference positions are read from smem tables; partial 3x3
// tensors are combined through smem.
template <int KP, int MT, int STAGES, bool WRAP, bool SELECT>
__global__ void __launch_bounds__(kAffineWarps * 32, (KP <= 9 && MT == 1) ? 2 : 1)
    affine_tma_kernel(const double* __restrict__ in, const double* __restrict__ ref, const double* __restrict__ G,
                      int64_t num_frames, int K, int row_stride, int stage_doubles, Alpha0 a0,
                      double* __restrict__ alpha, AlphaPeers peers, int peers_bulk, TileSelect sel) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWS = 8 * MT;
    constexpr int RED = kAffineWarps * ROWS * kRedStride;  // doubles per reduction buffer
    double* stages = reinterpret_cast<double*>(smem_raw);
    const int slack = kAffineWarps * 8 * KP;
    double* red = stages + STAGES * stage_doubles + slack;
    double* tab_ref = red + 2 * RED;       // [kAffineWarps*8*KP] wrapped reference positions
    double* tab_g9 = tab_ref + slack;      // [kAffineWarps*8*KP] column 8 of G
    // [2][ROWS*9] finished rows staged for the bulk stores to the peers.  (A deeper ring — 8 or 16 tiles with
    // cp.async.bulk.wait_group.read 6 / 14 — was measured on 4 GPUs: 0.87 / 0.95 ms against 0.85 ms; more bulk
    // stores in flight per CTA make the kernel slower, not more tolerant of NVLink bursts.)
    double* outbuf = tab_g9 + slack;
    uint64_t* bars = reinterpret_cast<uint64_t*>(outbuf + 2 * ROWS * 9);
    const uint32_t full0 = smem_u32(bars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the tiles of this launch: all of them, or the selection of one phase of the pipelined schedule (then
    // num_frames is a multiple of ROWS).  Two cursors walk this CTA's tiles: the one computed on, and the one
    // being fetched STAGES visits ahead.
    // (SELECT is a template parameter: the cursor arithmetic in the tile loop costs the plain launch 3-4 %)
    const int64_t num_tiles = SELECT ? sel.count : (num_frames + ROWS - 1) / ROWS;
    const uint32_t row_bytes = (uint32_t)K * 8u;

    auto issue_tile = [&](int64_t tile, int s) {
        const int64_t frame0 = (SELECT ? select_tile(sel, tile) : tile) * ROWS;  // (prologue only: a division per call)
        const int rows = (int)min((int64_t)ROWS, num_frames - frame0);
        mbar_arrive_expect_tx(full0 + 8 * s, row_bytes * rows);
        const uint32_t dst0 = smem_u32(stages + (size_t)s * stage_doubles);
        const double* src0 = in + frame0 * (int64_t)K;
        for (int r = 0; r < rows; r++)
            tma_bulk_g2s(dst0 + (uint32_t)(r * row_stride) * 8u, src0 + (int64_t)r * K, row_bytes, full0 + 8 * s);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int s = 0; s < STAGES; s++) {
            const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
            if (tile < num_tiles) issue_tile(tile, s);
        }
    }
    for (int e = threadIdx.x; e < slack; e += blockDim.x) {  // tables are zero-padded to `slack` rows
        tab_ref[e] = WRAP ? __ldg(ref + e) : 0.0;
        tab_g9[e] = __ldg(G + (int64_t)e * 9 + 8);
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int kbase = warp * 8 * KP + 2 * t;
    double b8[KP][2];
    uint32_t validmask = 0;
#pragma unroll
    for (int p = 0; p < KP; p++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = kbase + 8 * p + h;
            b8[p][h] = __ldg(G + (int64_t)e * 9 + g);
            if (e < K) validmask |= 1u << (2 * p + h);
        }
    }
    const double* my_ref = tab_ref + kbase;
    const double* my_g9 = tab_g9 + kbase;
    const int rowoff = g * row_stride + kbase;
    double* myred = red + ((size_t)warp * ROWS + g) * kRedStride;

    int64_t i = 0;
    int64_t pending_row0 = -1;  // first row of the tile that sits in outbuf waiting for its bulk stores
    TileCursor here, ahead;
    if (SELECT) {
        cursor_set(sel, blockIdx.x, here);
        cursor_set(sel, blockIdx.x + (int64_t)STAGES * gridDim.x, ahead);
    }
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, i++) {
        const int s = (int)(i % STAGES);
        const int64_t tile_row0 = (SELECT ? cursor_tile(sel, here) : tile) * ROWS;
        // full tiles go to the peers as bulk stores; a partial last tile uses plain stores
        const bool stage_rows = peers_bulk && (tile_row0 + ROWS <= num_frames);
        mbar_wait(full0 + 8 * s, (uint32_t)(i / STAGES) & 1);
        const double* base = stages + (size_t)s * stage_doubles + rowoff;
        // independent accumulator chains (2 per 8-frame group): one DMMA/DFMA dependency chain
        // per warp cannot cover the FP64 pipe latency with two warps per scheduler
        double c0[MT][2], c1[MT][2], a9[MT][2];
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int h = 0; h < 2; h++) c0[m][h] = c1[m][h] = a9[m][h] = 0.0;
#pragma unroll
        for (int p = 0; p < KP; p++) {
            const double2 g9 = *reinterpret_cast<const double2*>(my_g9 + 8 * p);
            double2 rf = make_double2(0.0, 0.0);
            if (WRAP) rf = *reinterpret_cast<const double2*>(my_ref + 8 * p);
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const double2 v = *reinterpret_cast<const double2*>(base + (size_t)m * 8 * row_stride + 8 * p);
                double a_lo = v.x, a_hi = v.y;
                if (WRAP) {
                    a_lo = wrap_disp(a_lo, rf.x);
                    a_hi = wrap_disp(a_hi, rf.y);
                }
                a_lo = (validmask >> (2 * p)) & 1u ? a_lo : 0.0;
                a_hi = (validmask >> (2 * p + 1)) & 1u ? a_hi : 0.0;
                dmma884(c0[m][0], c1[m][0], a_lo, b8[p][0]);
                a9[m][0] = fma(a_lo, g9.x, a9[m][0]);
                dmma884(c0[m][1], c1[m][1], a_hi, b8[p][1]);
                a9[m][1] = fma(a_hi, g9.y, a9[m][1]);
            }
        }
        double* slot = myred + (size_t)(i & 1) * RED;
#pragma unroll
        for (int m = 0; m < MT; m++) {
            double a9s = a9[m][0] + a9[m][1];
            a9s += __shfl_xor_sync(0xffffffffu, a9s, 1);
            a9s += __shfl_xor_sync(0xffffffffu, a9s, 2);
            *reinterpret_cast<double2*>(slot + (size_t)m * 8 * kRedStride + 2 * t) =
                make_double2(c0[m][0] + c0[m][1], c1[m][0] + c1[m][1]);
            if (t == 0) slot[(size_t)m * 8 * kRedStride + 8] = a9s;
        }
        __syncthreads();  // partials visible; every warp is done reading stage s
        if (peers_bulk && threadIdx.x == 32 && pending_row0 >= 0) {
            // fused all-gather: the previous tile's finished rows (staged in smem, complete since this
            // barrier) go to every peer GPU as one TMA bulk store each (UBLKCP S2G over NVLink)
            const uint32_t src = smem_u32(outbuf + (size_t)((i - 1) & 1) * ROWS * 9);
            const uint32_t mask = alpha_peer_mask(peers, pending_row0, ROWS);
            const int64_t off = alpha_peer_offset(peers, pending_row0);
            for (int p = 0; p < peers.count; p++)
                if (((mask >> p) & 1u) && peers.ptr[p])
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(peers.ptr[p] + off),
                                 "r"(src), "r"((uint32_t)(ROWS * 72))
                                 : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        {
            // refill stage s: thread 0 arms the barrier, lane 0 of warp w copies rows w, w+8, ...
            const int64_t next = tile + (int64_t)STAGES * gridDim.x;
            if (next < num_tiles && lane == 0) {
                const int64_t frame0 = (SELECT ? cursor_tile(sel, ahead) : next) * ROWS;
                const int rows = (int)min((int64_t)ROWS, num_frames - frame0);
                if (warp == 0) mbar_arrive_expect_tx(full0 + 8 * s, row_bytes * rows);
                const uint32_t dst0 = smem_u32(stages + (size_t)s * stage_doubles);
                const double* src0 = in + frame0 * (int64_t)K;
                for (int r = warp; r < rows; r += kAffineWarps)
                    tma_bulk_g2s(dst0 + (uint32_t)(r * row_stride) * 8u, src0 + (int64_t)r * K, row_bytes,
                                 full0 + 8 * s);
            }
        }
        if (threadIdx.x < 9 * ROWS) {
            const int f = threadIdx.x / 9, q = threadIdx.x % 9;
            const double* r = red + (size_t)(i & 1) * RED + (size_t)f * kRedStride + q;
            double sum = 0;
#pragma unroll
            for (int w = 0; w < kAffineWarps; w++) sum += r[(size_t)w * ROWS * kRedStride];
            const int64_t frame = tile_row0 + f;
            const double value = sum + a0.v[q];
            if (frame < num_frames) alpha[frame * 9 + q] = value;
            if (stage_rows) {
                outbuf[(size_t)(i & 1) * ROWS * 9 + threadIdx.x] = value;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async-proxy reads
            } else if (frame < num_frames) {
                // partial last tile / unaligned peers: plain stores over NVLink
                const uint32_t mask = alpha_peer_mask(peers, frame, 1);
                const int64_t off = alpha_peer_offset(peers, frame);
                for (int p = 0; p < peers.count; p++)
                    if (((mask >> p) & 1u) && peers.ptr[p]) peers.ptr[p][off + q] = value;
            }
        }
        if (peers_bulk && threadIdx.x == 32) {
            // the staging buffer written two tiles from now must not be read by a bulk store any more
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        pending_row0 = stage_rows ? tile_row0 : -1;
        if (SELECT) {
            cursor_advance(sel, gridDim.x, here);
            cursor_advance(sel, gridDim.x, ahead);
        }
    }
    if (peers_bulk) {
        __syncthreads();
        if (threadIdx.x == 32 && pending_row0 >= 0) {
            const uint32_t src = smem_u32(outbuf + (size_t)((i - 1) & 1) * ROWS * 9);
            const uint32_t mask = alpha_peer_mask(peers, pending_row0, ROWS);
            const int64_t off = alpha_peer_offset(peers, pending_row0);
            for (int p = 0; p < peers.count; p++)
                if (((mask >> p) & 1u) && peers.ptr[p])
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(peers.ptr[p] + off),
                                 "r"(src), "r"((uint32_t)(ROWS * 72))
                                 : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (threadIdx.x == 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static std::atomic<int> g_affine_mt{0};  // 0 = automatic (A/B switch, include/ramannoodle_b200_debug.h)

template <int KP, int MT, int STAGES, bool WRAP>
static int launch_affine_tma_cfg(const rn_model* m, const double* d_in, int64_t frames, double* d_alpha, Alpha0 a0,
                                 cudaStream_t stream, const AlphaPeers& peers) {
    const int K = (int)m->dim;
    const AffineSmemLayout L = affine_layout(K, KP, MT, STAGES);
    if (L.bytes > 227 * 1024) return 1;  // does not fit: caller tries a smaller configuration
    const TileSelect sel = peers.sel_stripe > 0 ? phase_tiles(peers.first_frame, frames, peers.sel_stripe, peers.sel_phase, 8 * MT)
                                                : all_tiles();
    const int64_t tiles = sel.period ? sel.count : (frames + 8 * MT - 1) / (8 * MT);
    if (tiles == 0) return RN_OK;
    const double* ref = WRAP ? m->d_ref_wrapped : m->d_zero_ref;
    const double* G = WRAP ? m->d_g_frac : m->d_g_cart;
    auto kern = sel.period ? affine_tma_kernel<KP, MT, STAGES, WRAP, true> : affine_tma_kernel<KP, MT, STAGES, WRAP, false>;
    RN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
    int ctas_per_sm = 1;
    RN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kAffineWarps * 32, L.bytes));
    if (ctas_per_sm < 1) return 1;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)m->sm_count * ctas_per_sm);
    // bulk (TMA) stores to the peers need 16-byte aligned destinations
    int peers_bulk = peers.count > 0 ? 1 : 0;
    for (int p = 0; p < peers.count; p++)
        if (peers.ptr[p] && reinterpret_cast<uintptr_t>(peers.ptr[p]) % 16 != 0) peers_bulk = 0;
    if (peers.log2_period >= 0 && (peers.first_frame * 72) % 16 != 0) peers_bulk = 0;
    // (plain st.global to the peers instead of bulk stores: 0.96 ms against 0.85 ms on 4 GPUs)
    kern<<<grid, kAffineWarps * 32, L.bytes, stream>>>(d_in, ref, G, frames, K, L.row_stride, L.stage_doubles, a0,
                                                       d_alpha, peers, peers_bulk, sel);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

template <int KP, bool WRAP>
static int launch_affine_tma_kp(const rn_model* m, const double* d_in, int64_t frames, double* d_alpha, Alpha0 a0,
                                cudaStream_t stream, const AlphaPeers& peers) {
    // Measured on B200 (tools/tune_affine.py, 1M frames): LLZO (KP=9) 16-frame tiles, 1 CTA/SM:
    // 6284 GB/s; 8-frame tiles, 2 CTAs/SM: 6200 GB/s.  TiO2 (KP=6): 4855 vs 5127 GB/s.
    int mt = g_affine_mt.load(std::memory_order_relaxed);
    if (mt == 0) mt = (KP >= 8) ? 2 : 1;
    int rc = 1;
    if (mt == 2) rc = launch_affine_tma_cfg<KP, 2, 2, WRAP>(m, d_in, frames, d_alpha, a0, stream, peers);
    if (rc == 1) rc = launch_affine_tma_cfg<KP, 1, 2, WRAP>(m, d_in, frames, d_alpha, a0, stream, peers);
    return rc;
}

template <bool WRAP>
static int launch_affine_tma(const rn_model* m, const double* d_in, int64_t frames, double* d_alpha, Alpha0 a0,
                             cudaStream_t stream, const AlphaPeers& peers) {
    switch (m->affine_kp) {
#define RN_AFFINE_CASE(KP) \
    case KP: return launch_affine_tma_kp<KP, WRAP>(m, d_in, frames, d_alpha, a0, stream, peers);
        RN_AFFINE_CASE(1)
        RN_AFFINE_CASE(2)
        RN_AFFINE_CASE(3)
        RN_AFFINE_CASE(4)
        RN_AFFINE_CASE(5)
        RN_AFFINE_CASE(6)
        RN_AFFINE_CASE(7)
        RN_AFFINE_CASE(8)
        RN_AFFINE_CASE(9)
        RN_AFFINE_CASE(10)
        RN_AFFINE_CASE(11)
        RN_AFFINE_CASE(12)
#undef RN_AFFINE_CASE
        default:
            set_error("affine TMA kernel not compiled for KP=%d", m->affine_kp);
            return RN_ERR_UNSUPPORTED;
    }
}

template <bool WRAP>
static int launch_affine_generic(const rn_model* m, const double* d_in, int64_t begin, int64_t end, double* d_alpha,
                                 Alpha0 a0, cudaStream_t stream) {
    if (end <= begin) return RN_OK;
    const int64_t groups = (end - begin + 7) / 8;
    const int grid = (int)std::min<int64_t>((groups + 7) / 8, (int64_t)m->sm_count * 8);
    affine_generic_kernel<WRAP><<<grid, 256, 0, stream>>>(d_in, WRAP ? m->d_ref_wrapped : m->d_zero_ref,
                                                          WRAP ? m->d_g_frac : m->d_g_cart, begin, end, (int)m->dim, a0,
                                                          d_alpha);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

static std::atomic<bool> g_force_generic_affine{false};

int launch_affine(const rn_model* m, const double* d_in, bool wrap, int64_t num_frames, double* d_alpha,
                  cudaStream_t stream, const AlphaPeers* peers, bool* peers_done) {
    Alpha0 a0;
    for (int q = 0; q < 9; q++) a0.v[q] = m->alpha0[q];
    const int K = (int)m->dim;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_in) % 16) == 0 && (K % 2 == 0);
    int64_t tma_frames = 0;
    if (m->affine_kp > 0 && aligned && !g_force_generic_affine.load(std::memory_order_relaxed)) {
        tma_frames = num_frames;
    }
    int rc = RN_OK;
    AlphaPeers fused = no_peers();
    if (peers) fused = *peers;
    if (tma_frames > 0) {
        rc = wrap ? launch_affine_tma<true>(m, d_in, tma_frames, d_alpha, a0, stream, fused)
                  : launch_affine_tma<false>(m, d_in, tma_frames, d_alpha, a0, stream, fused);
        if (rc == 1) {  // no configuration fits in shared memory
            tma_frames = 0;
            rc = RN_OK;
        }
        if (rc != RN_OK) return rc;
    }
    if (fused.sel_stripe > 0 && tma_frames != num_frames) {
        set_error("a frame selection (pipelined multi-GPU schedule) needs the TMA affine kernel");
        return RN_ERR_UNSUPPORTED;
    }
    // the TMA kernel stores to the peers itself; the generic fallback does not
    if (peers_done) *peers_done = (tma_frames == num_frames);
    if (tma_frames < num_frames) {
        rc = wrap ? launch_affine_generic<true>(m, d_in, tma_frames, num_frames, d_alpha, a0, stream)
                  : launch_affine_generic<false>(m, d_in, tma_frames, num_frames, d_alpha, a0, stream);
    }
    return rc;
}

// ------------------------------------------------------------------------------------
// alpha = alpha0 for models without DOFs
// ------------------------------------------------------------------------------------
__global__ void fill_alpha0_kernel(int64_t count, Alpha0 a0, double* __restrict__ alpha) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        alpha[i] = a0.v[i % 9];
}

int launch_fill_alpha0(const rn_model* m, int64_t num_frames, double* d_alpha, cudaStream_t stream) {
    if (num_frames == 0) return RN_OK;
    Alpha0 a0;
    for (int q = 0; q < 9; q++) a0.v[q] = m->alpha0[q];
    const int64_t count = num_frames * 9;
    const int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)m->sm_count * 8);
    fill_alpha0_kernel<<<grid, 256, 0, stream>>>(count, a0, d_alpha);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

// ------------------------------------------------------------------------------------
// dense kernel: DMMA projection + fused spline epilogue
// ------------------------------------------------------------------------------------
constexpr int kFT = 128;  // frames per CTA tile (8 warps x 16)
constexpr int kJT = 64;   // DOFs per J tile
constexpr int kKC = 16;   // K elements per pipeline chunk
constexpr int kRS = 20;   // padded smem row stride (doubles): conflict-free fragment loads
constexpr int kDStages = 3;

template <int DEG, bool WRAP, bool ALIGN16>
__global__ void __launch_bounds__(256, 1)
    dense_kernel(const double* __restrict__ in, const double* __restrict__ ref, const double* __restrict__ V,
                 const int32_t* __restrict__ piece_off, const double* __restrict__ breaks,
                 const double* __restrict__ pieces, int64_t num_frames, int K, int Kv, int Jpad, int accumulate,
                 Alpha0 a0, double* __restrict__ alpha) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);      // [kDStages][kFT][kRS]
    double* Bs = As + (size_t)kDStages * kFT * kRS;        // [kDStages][kJT][kRS]
    constexpr int REC = 1 + 9 * (DEG + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int chunks = Kv / kKC;
    const int jtiles = Jpad / kJT;
    const int64_t total = (int64_t)jtiles * chunks;
    const int64_t num_tiles = (num_frames + kFT - 1) / kFT;

    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t frame0 = tile * kFT;

        auto issue = [&](int64_t it) {
            if (it < total) {
                const int jt = (int)(it / chunks), kc = (int)(it % chunks);
                const int st = (int)(it % kDStages);
                double* a_dst = As + (size_t)st * kFT * kRS;
                double* b_dst = Bs + (size_t)st * kJT * kRS;
                if (ALIGN16) {
#pragma unroll
                    for (int r = 0; r < (kFT * 8) / 256; r++) {
                        const int idx = threadIdx.x + 256 * r;
                        const int row = idx >> 3, seg = idx & 7;
                        const int col = kc * kKC + seg * 2;
                        const int64_t frame = frame0 + row;
                        int bytes = 0;
                        if (frame < num_frames && col < K) bytes = min(16, (K - col) * 8);
                        const double* src = in + (bytes ? frame * (int64_t)K + col : 0);
                        cp_async_16(smem_u32(a_dst + row * kRS + seg * 2), src, bytes);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < (kFT * 16) / 256; r++) {
                        const int idx = threadIdx.x + 256 * r;
                        const int row = idx >> 4, seg = idx & 15;
                        const int col = kc * kKC + seg;
                        const int64_t frame = frame0 + row;
                        const int bytes = (frame < num_frames && col < K) ? 8 : 0;
                        const double* src = in + (bytes ? frame * (int64_t)K + col : 0);
                        cp_async_8(smem_u32(a_dst + row * kRS + seg), src, bytes);
                    }
                }
#pragma unroll
                for (int r = 0; r < (kJT * 8) / 256; r++) {
                    const int idx = threadIdx.x + 256 * r;
                    const int row = idx >> 3, seg = idx & 7;
                    const double* src = V + ((int64_t)jt * kJT + row) * Kv + kc * kKC + seg * 2;
                    cp_async_16(smem_u32(b_dst + row * kRS + seg * 2), src, 16);
                }
            }
            cp_async_commit();
        };

        double out9[2][9];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int q = 0; q < 9; q++) out9[mt][q] = 0.0;
        double acc[2][8][2];

        __syncthreads();  // previous tile's smem reads are finished
        for (int p = 0; p < kDStages - 1; p++) issue(p);

        for (int64_t it = 0; it < total; it++) {
            const int kc = (int)(it % chunks);
            if (kc == 0) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            }
            cp_async_wait<kDStages - 2>();
            __syncthreads();
            issue(it + kDStages - 1);
            const int st = (int)(it % kDStages);
            const double* a_src = As + (size_t)st * kFT * kRS + (warp * 16 + g) * kRS + t;
            const double* b_src = Bs + (size_t)st * kJT * kRS + g * kRS + t;
#pragma unroll
            for (int s = 0; s < kKC / 4; s++) {
                double a[2];
                a[0] = a_src[4 * s];
                a[1] = a_src[8 * kRS + 4 * s];
                if (WRAP) {
                    const double r = __ldg(ref + kc * kKC + 4 * s + t);
                    a[0] = wrap_disp(a[0], r);
                    a[1] = wrap_disp(a[1], r);
                }
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const double b = b_src[nt * 8 * kRS + 4 * s];
                    dmma884(acc[0][nt][0], acc[0][nt][1], a[0], b);
                    dmma884(acc[1][nt][0], acc[1][nt][1], a[1], b);
                }
            }
            if (kc == chunks - 1) {
                // ---- fused epilogue: amplitudes -> piecewise polynomials -> 3x3 partial sums ----
                const int jt = (int)(it / chunks);
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int j = jt * kJT + nt * 8 + 2 * t + c;
                        const int p0 = __ldg(piece_off + j);
                        const int np = __ldg(piece_off + j + 1) - p0;
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) {
                            const double x = acc[mt][nt][c];
                            int p = 0;
                            for (int b = 0; b + 1 < np; b++) p += (x >= __ldg(breaks + p0 + b)) ? 1 : 0;
                            const double* rec = pieces + (int64_t)(p0 + p) * REC;
                            const double dx = x - __ldg(rec);
#pragma unroll
                            for (int q = 0; q < 9; q++) {
                                double r = __ldg(rec + 1 + q);
#pragma unroll
                                for (int mm = 1; mm <= DEG; mm++) r = fma(r, dx, __ldg(rec + 1 + 9 * mm + q));
                                out9[mt][q] += r;
                            }
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();

        // ---- reduce the 4 lanes that share a frame, then store ----
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const int64_t frame = frame0 + warp * 16 + mt * 8 + g;
#pragma unroll
            for (int q = 0; q < 9; q++) {
                double v = out9[mt][q];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                out9[mt][q] = v;
            }
            if (frame < num_frames) {
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    if ((q & 3) == t) {
                        const double base = accumulate ? alpha[frame * 9 + q] : a0.v[q];
                        alpha[frame * 9 + q] = out9[mt][q] + base;
                    }
                }
            }
        }
    }
}

template <int DEG>
static int launch_dense_deg(const rn_model* m, const double* d_in, bool wrap, bool accumulate, int64_t num_frames,
                            double* d_alpha, cudaStream_t stream) {
    Alpha0 a0;
    for (int q = 0; q < 9; q++) a0.v[q] = m->alpha0[q];
    const int K = (int)m->dim;
    const bool align16 = (reinterpret_cast<uintptr_t>(d_in) % 16 == 0) && (K % 2 == 0);
    const size_t smem = (size_t)kDStages * (kFT + kJT) * kRS * sizeof(double);
    const int64_t tiles = (num_frames + kFT - 1) / kFT;
    const int grid = (int)std::min<int64_t>(tiles, m->sm_count);
    const double* ref = m->d_ref_wrapped;
    const double* V = wrap ? m->d_v_frac : m->d_v_cart;
#define RN_DENSE_LAUNCH(W, A)                                                                                  \
    {                                                                                                          \
        auto kern = dense_kernel<DEG, W, A>;                                                                   \
        RN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<grid, 256, smem, stream>>>(d_in, ref, V, m->d_piece_off, m->d_breaks, m->d_pieces, num_frames, K, \
                                          (int)m->v_cols, (int)m->dense_pad, accumulate ? 1 : 0, a0, d_alpha); \
    }
    if (wrap) {
        if (align16) RN_DENSE_LAUNCH(true, true) else RN_DENSE_LAUNCH(true, false)
    } else {
        if (align16) RN_DENSE_LAUNCH(false, true) else RN_DENSE_LAUNCH(false, false)
    }
#undef RN_DENSE_LAUNCH
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}

int launch_dense_v1(const rn_model* m, const double* d_in, bool wrap, bool accumulate, int64_t num_frames,
                    double* d_alpha, cudaStream_t stream) {
    if (num_frames == 0) return RN_OK;
    switch (m->dense_degree) {
        case 0:
        case 1: return launch_dense_deg<1>(m, d_in, wrap, accumulate, num_frames, d_alpha, stream);
        case 2: return launch_dense_deg<2>(m, d_in, wrap, accumulate, num_frames, d_alpha, stream);
        case 3: return launch_dense_deg<3>(m, d_in, wrap, accumulate, num_frames, d_alpha, stream);
        case 4: return launch_dense_deg<4>(m, d_in, wrap, accumulate, num_frames, d_alpha, stream);
        case 5: return launch_dense_deg<5>(m, d_in, wrap, accumulate, num_frames, d_alpha, stream);
        default:
            set_error("dense kernel not compiled for spline degree %d", m->dense_degree);
            return RN_ERR_UNSUPPORTED;
    }
}

}  // namespace rn

using namespace rn;

static int eval_common(const rn_model* model, const double* d_in, bool wrap, int64_t num_frames, double* d_alpha,
                       void* stream, const AlphaPeers* peers = nullptr) {
    RN_CHECK_ARG(model != nullptr, "model is null");
    RN_CHECK_ARG(num_frames >= 0, "num_frames must be non-negative");
    if (num_frames == 0) return RN_OK;
    RN_CHECK_ARG(d_in && d_alpha, "null device pointer");
    DeviceGuard guard(model->device);
    if (!guard.ok) {
        set_error("cudaSetDevice(%d) failed", model->device);
        return RN_ERR_CUDA;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (peers && peers->count == 0) peers = nullptr;
    int rc = RN_OK;
    bool peers_done = false;
    if (model->num_dofs == 0) {
        rc = launch_fill_alpha0(model, num_frames, d_alpha, s);
    } else {
        const bool run_affine = model->num_linear > 0;
        const bool run_dense = model->num_dense > 0;
        if (run_affine) {
            // if a dense pass follows, the final rows come from it: fuse the peer stores there
            rc = launch_affine(model, d_in, wrap, num_frames, d_alpha, s, run_dense ? nullptr : peers, &peers_done);
            if (rc != RN_OK) return rc;
        }
        if (run_dense) rc = launch_dense(model, d_in, wrap, run_affine, num_frames, d_alpha, s, peers, &peers_done);
    }
    if (rc != RN_OK) return rc;
    if (peers && !peers_done) {
        // kernels without fused peer stores: plain device-to-peer copies of the whole block (routed mode:
        // to every rank — rows a rank does not own are simply never read there)
        for (int p = 0; p < peers->count; p++)
            if (peers->ptr[p])
                RN_CUDA(cudaMemcpyAsync(peers->ptr[p] + alpha_peer_offset(*peers, 0), d_alpha,
                                        sizeof(double) * 9 * num_frames, cudaMemcpyDeviceToDevice, s));
    }
    return RN_OK;
}

static int make_peers(double* const* d_alpha_outputs, int num_outputs, AlphaPeers* peers) {
    RN_CHECK_ARG(d_alpha_outputs != nullptr && num_outputs >= 1 && num_outputs <= 8,
                 "between 1 and 8 output pointers are required");
    *peers = no_peers();
    peers->count = num_outputs - 1;
    for (int i = 0; i < num_outputs; i++) RN_CHECK_ARG(d_alpha_outputs[i] != nullptr, "null output pointer");
    for (int i = 1; i < num_outputs; i++) peers->ptr[i - 1] = d_alpha_outputs[i];
    return RN_OK;
}

extern "C" int rn_calc_polarizabilities_multi(const rn_model* model, const double* d_positions, int64_t num_frames,
                                              double* const* d_alpha_outputs, int num_outputs, void* stream) {
    AlphaPeers peers;
    int rc = make_peers(d_alpha_outputs, num_outputs, &peers);
    if (rc != RN_OK) return rc;
    return eval_common(model, d_positions, true, num_frames, d_alpha_outputs[0], stream, &peers);
}

int rn::make_routed_peers(double* const* peer_series, int world, int64_t first_frame, int64_t period, int64_t width,
                          AlphaPeers* peers) {
    RN_CHECK_ARG(peer_series != nullptr && world >= 1 && world <= 8, "between 1 and 8 ranks");
    RN_CHECK_ARG(first_frame >= 0, "first_frame must be non-negative");
    RN_CHECK_ARG(period > 0 && (period & (period - 1)) == 0 && width >= 32 && (width & (width - 1)) == 0 &&
                     width <= period && period / width <= world,
                 "period and width must be powers of two with width >= 32 and period / width <= world");
    *peers = no_peers();
    peers->count = world;
    for (int r = 0; r < world; r++) peers->ptr[r] = peer_series[r];
    peers->log2_period = 0;
    while (((int64_t)1 << peers->log2_period) < period) peers->log2_period++;
    peers->log2_width = 0;
    while (((int64_t)1 << peers->log2_width) < width) peers->log2_width++;
    peers->first_frame = first_frame;
    return RN_OK;
}

int rn::eval_with_peers(const rn_model* model, const double* d_positions, int64_t num_frames, double* d_alpha,
                        cudaStream_t stream, const AlphaPeers& peers) {
    return eval_common(model, d_positions, true, num_frames, d_alpha, stream, &peers);
}

extern "C" int rn_calc_polarizabilities_routed(const rn_model* model, const double* d_positions, int64_t num_frames,
                                               double* d_alpha, double* const* peer_series, int world,
                                               int64_t first_frame, int64_t period, int64_t width, void* stream) {
    AlphaPeers peers;
    int rc = make_routed_peers(peer_series, world, first_frame, period, width, &peers);
    if (rc != RN_OK) return rc;
    return eval_common(model, d_positions, true, num_frames, d_alpha, stream, &peers);
}

// One phase of the pipelined schedule: the frames n of the block with (n mod 2 stripe) < stripe + 16 (phase 0:
// everything the packs of the stripes' first halves read, their rows n + 1 included) or the others (phase 1).
static bool routed_phases_ok(const rn_model* model, const double* d_positions, int64_t num_frames, int64_t first_frame,
                             int64_t stripe) {
    if (!model || model->num_dofs == 0 || model->num_dense > 0 || model->num_linear == 0 || model->affine_kp <= 0) return false;
    if (g_force_generic_affine.load(std::memory_order_relaxed)) return false;
    if (reinterpret_cast<uintptr_t>(d_positions) % 16 != 0 || model->dim % 2 != 0) return false;
    return stripe >= 1024 && (stripe & (stripe - 1)) == 0 && num_frames % 16 == 0 && first_frame % 16 == 0 &&
           num_frames >= 16;
}

extern "C" int rn_routed_phases_supported(const rn_model* model, const double* d_positions, int64_t num_frames,
                                          int64_t first_frame, int64_t stripe) {
    return routed_phases_ok(model, d_positions, num_frames, first_frame, stripe) ? 1 : 0;
}

extern "C" int rn_calc_polarizabilities_routed_phase(const rn_model* model, const double* d_positions,
                                                     int64_t num_frames, double* d_alpha, double* const* peer_series,
                                                     int world, int64_t first_frame, int64_t period, int64_t width,
                                                     int64_t stripe, int phase, void* stream) {
    RN_CHECK_ARG(phase == 0 || phase == 1, "phase must be 0 or 1");
    if (!routed_phases_ok(model, d_positions, num_frames, first_frame, stripe)) {
        set_error("this model / trajectory block cannot be evaluated in phases (rn_routed_phases_supported)");
        return RN_ERR_UNSUPPORTED;
    }
    AlphaPeers peers;
    int rc = make_routed_peers(peer_series, world, first_frame, period, width, &peers);
    if (rc != RN_OK) return rc;
    peers.sel_stripe = stripe;
    peers.sel_phase = phase;
    return eval_common(model, d_positions, true, num_frames, d_alpha, stream, &peers);
}

extern "C" int rn_calc_polarizabilities(const rn_model* model, const double* d_positions, int64_t num_frames,
                                        double* d_alpha, void* stream) {
    return eval_common(model, d_positions, true, num_frames, d_alpha, stream);
}

extern "C" int rn_get_polarizability(const rn_model* model, const double* d_cart_displacements, int64_t num_frames,
                                     double* d_alpha, void* stream) {
    return eval_common(model, d_cart_displacements, false, num_frames, d_alpha, stream);
}

// Test hook: the local 16-frame tiles one phase of rn_calc_polarizabilities_routed_phase evaluates, in launch order
// (host arithmetic only; returns their number, fills at most `capacity`).
extern "C" int64_t rn_debug_phase_tiles(int64_t num_frames, int64_t first_frame, int64_t stripe, int phase,
                                        int64_t* tiles, int64_t capacity) {
    // (walked with the cursor the kernel uses, from two starting points like two CTAs of a grid of 2)
    const TileSelect sel = phase_tiles(first_frame, num_frames, stripe, phase, 16);
    for (int64_t start = 0; start < 2; start++) {
        TileCursor c;
        cursor_set(sel, start, c);
        for (int64_t j = start; j < sel.count; j += 2, cursor_advance(sel, 2, c)) {
            if (cursor_tile(sel, c) != select_tile(sel, j)) return -1;
            if (j < capacity) tiles[j] = cursor_tile(sel, c);
        }
    }
    return sel.count;
}

// Test hook (include/ramannoodle_b200_debug.h): route the affine term through the generic kernel.
extern "C" void rn_debug_force_generic_affine(int on) { rn::g_force_generic_affine.store(on != 0, std::memory_order_relaxed); }
// Tuning hook: frames-per-tile multiplier (1 or 2) and pipeline depth of the TMA affine kernel.
extern "C" void rn_debug_set_affine_config(int mt, int stages) {
    (void)stages;
    rn::g_affine_mt.store(mt, std::memory_order_relaxed);
}

__global__ void apply_pbc_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const double p = in[i];
        out[i] = p - floor(p);  // positions - positions // 1  (structure/utils.py:27)
    }
}

extern "C" int rn_apply_pbc(const double* d_in, double* d_out, int64_t count, void* stream) {
    RN_CHECK_ARG(count >= 0, "count must be non-negative");
    if (count == 0) return RN_OK;
    RN_CHECK_ARG(d_in && d_out, "null device pointer");
    const int grid = (int)std::min<int64_t>((count + 255) / 256, 148 * 16);
    apply_pbc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, d_out, count);
    RN_LAUNCHED();
    RN_CUDA(cudaGetLastError());
    return RN_OK;
}
