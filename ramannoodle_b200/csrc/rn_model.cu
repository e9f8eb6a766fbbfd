// Model packing: turns the reference's evaluation state (basis vectors, B-splines, mask)
// into the device tables the kernels read.  Host-side work, done once per model.
//
// What is precomputed (all in 80-bit long double, rounded once to fp64):
//   * every B-spline is converted to a piecewise polynomial in local powers of (x - x0),
//     piece selection reproducing scipy's find_interval (extrapolate=True);
//   * DOFs whose interpolant is ONE linear piece (every ARTModel DOF: knots [-a,-a,a,a],
//     ramannoodle/pmodel/_art.py:187-197) are collapsed into the affine term
//     alpha = alpha0 + D . G, G = sum_j w_j v_j (x) slope_j   (3N x 9);
//   * the lattice multiply of get_cart_displacement (structure/_reference.py:285) is folded
//     into G and into the dense basis V, so kernels contract wrapped *fractional*
//     displacements directly.
// Re-association only: measured head-room vs the reference is ~1e-14 relative.
#include <cmath>
#include <cstring>
#include <limits>

#include "rn_common.cuh"

namespace rn {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    g_last_error = buf;
}

using LD = long double;

struct PiecewisePoly {
    int degree = 0;
    std::vector<double> breaks;  // pieces-1 thresholds: piece p is used when breaks[p-1] <= x (< breaks[p])
    std::vector<double> x0;      // pieces
    std::vector<LD> coefs;       // pieces * (degree+1) * 9, index ((p*(degree+1) + m)*9 + q), power m
    int pieces() const { return static_cast<int>(x0.size()); }
};

// B-spline (t, c (n,9), k) -> piecewise polynomial.  Mirrors scipy's evaluate_spline:
// find_interval picks ell = max({k} U {l in [k+1, n-1] : t[l] <= x}); _deBoor_D builds the
// k+1 non-zero basis functions on [t[ell], t[ell+1]) with zero-width spans skipped.  Here the
// same recursion is run on polynomials in u = x - t[ell].
static int bspline_to_pp(const double* t, int nt, const double* c, int k, PiecewisePoly& pp) {
    const int n = nt - k - 1;
    if (k < 0 || k > kMaxDegree) {
        set_error("unsupported spline degree %d (supported: 0..%d)", k, kMaxDegree);
        return RN_ERR_UNSUPPORTED;
    }
    if (n < k + 1) {
        set_error("spline with %d knots and degree %d has too few coefficients", nt, k);
        return RN_ERR_INVALID_ARGUMENT;
    }
    for (int i = 1; i < nt; i++) {
        if (!(t[i] >= t[i - 1])) {
            set_error("spline knots must be non-decreasing and finite");
            return RN_ERR_INVALID_ARGUMENT;
        }
    }
    std::vector<int> ells;
    ells.push_back(k);
    for (int l = k + 1; l <= n - 1; l++) {
        if (ells.size() > 1 && t[l] == t[ells.back()]) {
            ells.back() = l;  // same threshold: the larger index wins
        } else {
            ells.push_back(l);
        }
    }
    pp.degree = k;
    pp.breaks.clear();
    pp.x0.clear();
    pp.coefs.assign(ells.size() * (size_t)(k + 1) * 9, 0.0L);
    for (size_t p = 0; p < ells.size(); p++) {
        const int ell = ells[p];
        if (p > 0) pp.breaks.push_back(t[ell]);
        const LD x0 = t[ell];
        pp.x0.push_back(t[ell]);
        LD h[kMaxDegree + 1][kMaxDegree + 1] = {{0}};
        LD hh[kMaxDegree + 1][kMaxDegree + 1] = {{0}};
        h[0][0] = 1.0L;
        for (int j = 1; j <= k; j++) {
            for (int a = 0; a < j; a++)
                for (int d = 0; d <= k; d++) hh[a][d] = h[a][d];
            for (int d = 0; d <= k; d++) h[0][d] = 0.0L;
            for (int nn = 1; nn <= j; nn++) {
                const double xb = t[ell + nn];
                const double xa = t[ell + nn - j];
                if (xb == xa) {
                    for (int d = 0; d <= k; d++) h[nn][d] = 0.0L;
                    continue;
                }
                const LD inv = 1.0L / ((LD)xb - (LD)xa);
                const LD A1 = (LD)xb - x0;  // (xb - x) = A1 - u
                const LD A2 = x0 - (LD)xa;  // (x - xa) = A2 + u
                for (int d = k; d >= 0; d--) {
                    const LD w_d = hh[nn - 1][d] * inv;
                    const LD w_dm1 = (d > 0) ? hh[nn - 1][d - 1] * inv : 0.0L;
                    h[nn - 1][d] += A1 * w_d - w_dm1;
                    h[nn][d] = A2 * w_d + w_dm1;
                }
            }
        }
        for (int a = 0; a <= k; a++) {
            const double* crow = c + (size_t)(ell + a - k) * 9;
            for (int m = 0; m <= k; m++)
                for (int q = 0; q < 9; q++) pp.coefs[((p * (k + 1)) + m) * 9 + q] += (LD)crow[q] * h[a][m];
        }
    }
    return RN_OK;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    RN_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
    RN_CUDA(cudaMemset(*dst, 0, bytes));
    if (!src.empty()) RN_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return RN_OK;
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static void destroy_model(rn_model* m) {
    if (!m) return;
    cudaFree(m->d_ref_wrapped);
    cudaFree(m->d_zero_ref);
    cudaFree(m->d_g_frac);
    cudaFree(m->d_g_cart);
    cudaFree(m->d_v_frac);
    cudaFree(m->d_v_cart);
    cudaFree(m->d_piece_off);
    cudaFree(m->d_breaks);
    cudaFree(m->d_pieces);
    cudaFree(m->d_tp_x0);
    cudaFree(m->d_tp_brk);
    cudaFree(m->d_tp_c8);
    cudaFree(m->d_tp_c9);
    delete m;
}

}  // namespace rn

using namespace rn;

extern "C" const char* rn_last_error(void) { return g_last_error.c_str(); }

extern "C" int64_t rn_launch_count(void) { return g_launch_count.load(); }

extern "C" int rn_device_info(int device, int* runtime_version, int* device_count, int* compute_capability,
                              int* sm_count) {
    int rt = 0, count = 0;
    RN_CUDA(cudaRuntimeGetVersion(&rt));
    RN_CUDA(cudaGetDeviceCount(&count));
    RN_CHECK_ARG(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
    cudaDeviceProp prop;
    RN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (runtime_version) *runtime_version = rt;
    if (device_count) *device_count = count;
    if (compute_capability) *compute_capability = prop.major * 10 + prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return RN_OK;
}

extern "C" int rn_bspline_to_pp(const double* h_knots, int num_knots, const double* h_coefs, int degree,
                                double* out_breaks, double* out_x0, double* out_coefs) {
    RN_CHECK_ARG(h_knots && h_coefs && out_breaks && out_x0 && out_coefs, "null pointer");
    PiecewisePoly pp;
    int rc = bspline_to_pp(h_knots, num_knots, h_coefs, degree, pp);
    if (rc != RN_OK) return rc;
    for (size_t i = 0; i < pp.breaks.size(); i++) out_breaks[i] = pp.breaks[i];
    for (size_t i = 0; i < pp.x0.size(); i++) out_x0[i] = pp.x0[i];
    for (size_t i = 0; i < pp.coefs.size(); i++) out_coefs[i] = (double)pp.coefs[i];
    return pp.pieces();
}

extern "C" int rn_model_create(const double* h_ref_positions, int64_t num_atoms, const double* h_lattice,
                               const double* h_basis, int64_t num_dofs, const int32_t* h_degree,
                               const int64_t* h_knot_off, const double* h_knots, const int64_t* h_coef_off,
                               const double* h_coefs, const double* h_weight,
                               const double* h_ref_polarizability, int device, int flags, rn_model** out) {
    RN_CHECK_ARG(out != nullptr, "out is null");
    *out = nullptr;
    RN_CHECK_ARG(h_ref_positions && h_lattice && h_ref_polarizability, "null structure pointer");
    RN_CHECK_ARG(num_atoms > 0, "num_atoms must be positive");
    RN_CHECK_ARG(num_dofs >= 0, "num_dofs must be non-negative");
    RN_CHECK_ARG(num_dofs == 0 || (h_basis && h_degree && h_knot_off && h_knots && h_coef_off && h_coefs && h_weight),
                 "null model table pointer");
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
    if (prop.major != 10) {
        set_error("ramannoodle_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                  prop.minor);
        return RN_ERR_UNSUPPORTED;
    }

    rn_model* m = new rn_model();
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->num_atoms = num_atoms;
    m->dim = 3 * num_atoms;
    m->num_dofs = num_dofs;
    const int64_t K = m->dim;
    const bool force_dense = (flags & RN_MODEL_FORCE_DENSE) != 0;
    {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&h](const double* data, size_t count) {
            const unsigned char* bytes = reinterpret_cast<const unsigned char*>(data);
            for (size_t i = 0; i < count * sizeof(double); i++) h = (h ^ bytes[i]) * 1099511628211ull;
        };
        mix(h_ref_positions, (size_t)K);
        mix(h_lattice, 9);
        m->ref_hash = h;
        // everything but the weights
        mix(h_ref_polarizability, 9);
        if (num_dofs > 0) {
            mix(h_basis, (size_t)num_dofs * (size_t)K);
            mix(h_knots, (size_t)h_knot_off[num_dofs]);
            mix(h_coefs, (size_t)h_coef_off[num_dofs] * 9);
            for (int64_t j = 0; j < num_dofs; j++) h = (h ^ (uint64_t)(h_degree[j] + 1)) * 1099511628211ull;
        }
        m->shape_hash = h ^ ((uint64_t)flags << 56);
    }

    // ---- splines -> piecewise polynomials; classify linear vs dense ----
    std::vector<PiecewisePoly> pps((size_t)num_dofs);
    std::vector<int64_t> dense_ids;
    std::vector<LD> alpha0(9);
    for (int q = 0; q < 9; q++) alpha0[q] = h_ref_polarizability[q];
    std::vector<LD> g_cart((size_t)K * 9, 0.0L);
    int dense_degree = 0, dense_max_pieces = 0;
    for (int64_t j = 0; j < num_dofs; j++) {
        const int64_t nt = h_knot_off[j + 1] - h_knot_off[j];
        const int64_t nc = h_coef_off[j + 1] - h_coef_off[j];
        const int k = h_degree[j];
        if (nt != nc + k + 1) {
            set_error("DOF %lld: %lld knots, %lld coefficients and degree %d are inconsistent", (long long)j,
                      (long long)nt, (long long)nc, k);
            destroy_model(m);
            return RN_ERR_INVALID_ARGUMENT;
        }
        int rc = bspline_to_pp(h_knots + h_knot_off[j], (int)nt, h_coefs + 9 * h_coef_off[j], k, pps[j]);
        if (rc != RN_OK) {
            destroy_model(m);
            return rc;
        }
        const PiecewisePoly& pp = pps[j];
        const LD w = h_weight[j];
        if (!force_dense && pp.pieces() == 1 && pp.degree <= 1) {
            // B(a) = c0 + c1 (a - x0): constant part into alpha0, slope into G
            const double* v = h_basis + j * K;
            for (int q = 0; q < 9; q++) {
                const LD c0 = pp.coefs[0 * 9 + q];
                const LD c1 = (pp.degree >= 1) ? pp.coefs[1 * 9 + q] : 0.0L;
                alpha0[q] += w * (c0 - c1 * (LD)pp.x0[0]);
                const LD s = w * c1;
                if (s != 0.0L) {
                    for (int64_t e = 0; e < K; e++) g_cart[(size_t)e * 9 + q] += s * (LD)v[e];
                }
            }
            m->num_linear++;
        } else {
            dense_ids.push_back(j);
            dense_degree = std::max(dense_degree, pp.degree);
            dense_max_pieces = std::max(dense_max_pieces, pp.pieces());
        }
    }
    m->num_dense = (int64_t)dense_ids.size();
    m->dense_degree = dense_degree;
    m->dense_max_pieces = dense_max_pieces;
    for (int q = 0; q < 9; q++) m->alpha0[q] = (double)alpha0[q];

    // ---- reference positions, wrapped once (structure/utils.py:132) ----
    const int kp = (int)((K + 8 * kAffineWarps - 1) / (8 * kAffineWarps));
    m->affine_kp = (kp <= kAffineMaxKP) ? kp : 0;
    m->g_rows = std::max<int64_t>(round_up(K, 8), (int64_t)kAffineWarps * 8 * kp);
    std::vector<double> ref((size_t)m->g_rows, 0.0), zero_ref((size_t)m->g_rows, 0.0);
    for (int64_t e = 0; e < K; e++) ref[e] = h_ref_positions[e] - std::floor(h_ref_positions[e]);

    // ---- affine tables ----
    std::vector<double> g_cart_d((size_t)m->g_rows * 9, 0.0), g_frac_d((size_t)m->g_rows * 9, 0.0);
    for (int64_t a = 0; a < num_atoms; a++) {
        for (int q = 0; q < 9; q++) {
            for (int cp = 0; cp < 3; cp++) {
                LD acc = 0.0L;
                for (int c = 0; c < 3; c++) acc += (LD)h_lattice[cp * 3 + c] * g_cart[(size_t)(a * 3 + c) * 9 + q];
                g_frac_d[(size_t)(a * 3 + cp) * 9 + q] = (double)acc;
                g_cart_d[(size_t)(a * 3 + cp) * 9 + q] = (double)g_cart[(size_t)(a * 3 + cp) * 9 + q];
            }
        }
    }

    // ---- dense tables ----
    const int64_t jd = m->num_dense;
    m->dense_pad = round_up(std::max<int64_t>(jd, 1), 64);
    m->v_cols = round_up(K, 16);
    std::vector<double> v_cart, v_frac, breaks, pieces;
    std::vector<int32_t> piece_off;
    const int rec = 1 + 9 * (dense_degree + 1);
    if (jd > 0) {
        v_cart.assign((size_t)m->dense_pad * m->v_cols, 0.0);
        v_frac.assign((size_t)m->dense_pad * m->v_cols, 0.0);
        piece_off.reserve((size_t)m->dense_pad + 1);
        for (int64_t i = 0; i < m->dense_pad; i++) {
            piece_off.push_back((int32_t)(pieces.size() / rec));
            if (i >= jd) {  // padding DOF: one all-zero piece
                pieces.insert(pieces.end(), (size_t)rec, 0.0);
                breaks.push_back(0.0);
                continue;
            }
            const int64_t j = dense_ids[(size_t)i];
            const double* v = h_basis + j * K;
            double* vc = v_cart.data() + (size_t)i * m->v_cols;
            double* vf = v_frac.data() + (size_t)i * m->v_cols;
            for (int64_t a = 0; a < num_atoms; a++) {
                for (int cp = 0; cp < 3; cp++) {
                    LD acc = 0.0L;
                    for (int c = 0; c < 3; c++) acc += (LD)h_lattice[cp * 3 + c] * (LD)v[a * 3 + c];
                    vf[a * 3 + cp] = (double)acc;
                    vc[a * 3 + cp] = v[a * 3 + cp];
                }
            }
            const PiecewisePoly& pp = pps[(size_t)j];
            const LD w = h_weight[j];
            for (int p = 0; p < pp.pieces(); p++) {
                pieces.push_back(pp.x0[(size_t)p]);
                // record order: highest power first (Horner), padded to dense_degree with zeros
                for (int mm = dense_degree; mm >= 0; mm--)
                    for (int q = 0; q < 9; q++)
                        pieces.push_back(mm <= pp.degree
                                             ? (double)(w * pp.coefs[((size_t)p * (pp.degree + 1) + mm) * 9 + q])
                                             : 0.0);
                breaks.push_back(p + 1 < pp.pieces() ? pp.breaks[(size_t)p] : 0.0);
            }
        }
        piece_off.push_back((int32_t)(pieces.size() / rec));
    }

    // ---- truncated-power tables for the chained-DMMA epilogue ----
    std::vector<double> tp_x0, tp_brk, tp_c8, tp_c9;
    if (jd > 0 && dense_degree >= 1 && dense_degree <= 3 && dense_max_pieces - 1 <= 3) {
        const int D = dense_degree;
        const int nbk = dense_max_pieces - 1;
        // jump polynomials at every interior break, in 80-bit arithmetic
        struct Jumps {
            std::vector<LD> d;  // (breaks, D+1, 9)
        };
        std::vector<Jumps> jumps((size_t)jd);
        bool spline_only = true;
        for (int64_t i = 0; i < jd; i++) {
            const PiecewisePoly& pp = pps[(size_t)dense_ids[(size_t)i]];
            const int P = pp.pieces(), dj = pp.degree;
            jumps[(size_t)i].d.assign((size_t)std::max(P - 1, 0) * (D + 1) * 9, 0.0L);
            LD scale = 0.0L;  // magnitude of the spline over its knot span
            const LD span = (P > 1) ? 2.0L * std::fabs((LD)pp.breaks.back() - (LD)pp.x0[0]) : 1.0L;  // ~ knot span
            for (int p = 0; p < P; p++)
                for (int mm = 0; mm <= dj; mm++)
                    for (int q = 0; q < 9; q++)
                        scale = std::max(scale, std::fabs(pp.coefs[((size_t)p * (dj + 1) + mm) * 9 + q]) * powl(span, mm));
            for (int b = 0; b + 1 < P; b++) {
                // Delta_b(u) = poly_{b+1}(u) - poly_b(u + h), h = x0_{b+1} - x0_b (binomial shift)
                const LD h = (LD)pp.x0[(size_t)b + 1] - (LD)pp.x0[(size_t)b];
                for (int q = 0; q < 9; q++) {
                    for (int mm = 0; mm <= dj; mm++) {
                        LD shifted = 0.0L;  // coefficient of u^mm in poly_b(u + h)
                        LD binom = 1.0L;    // C(r, mm) built incrementally over r
                        for (int r = mm; r <= dj; r++) {
                            if (r > mm) binom = binom * (LD)r / (LD)(r - mm);
                            shifted += pp.coefs[((size_t)b * (dj + 1) + r) * 9 + q] * binom * powl(h, r - mm);
                        }
                        const LD dv = pp.coefs[((size_t)(b + 1) * (dj + 1) + mm) * 9 + q] - shifted;
                        jumps[(size_t)i].d[((size_t)b * (D + 1) + mm) * 9 + q] = dv;
                        if (mm < D && std::fabs(dv) * powl(span, mm) > 1e-13L * scale) spline_only = false;
                    }
                }
            }
        }
        const int per_break = spline_only ? 1 : (D + 1);
        const int nf = D + nbk * per_break;
        if (nf <= 12) {
            m->tp_mode = spline_only ? 1 : 2;
            m->tp_breaks = nbk;
            m->tp_features = nf;
            const double inf = std::numeric_limits<double>::infinity();
            tp_x0.assign((size_t)m->dense_pad, 0.0);
            tp_brk.assign((size_t)m->dense_pad * std::max(nbk, 1), inf);
            tp_c8.assign((size_t)m->dense_pad * nf * 8, 0.0);
            tp_c9.assign((size_t)m->dense_pad * nf, 0.0);
            std::vector<LD> a0tp(9);
            for (int q = 0; q < 9; q++) a0tp[q] = alpha0[q];
            for (int64_t i = 0; i < jd; i++) {
                const int64_t j = dense_ids[(size_t)i];
                const PiecewisePoly& pp = pps[(size_t)j];
                const LD w = h_weight[j];
                const int dj = pp.degree, P = pp.pieces();
                tp_x0[(size_t)i] = pp.x0[0];
                for (int b = 0; b + 1 < P; b++) tp_brk[(size_t)i * std::max(nbk, 1) + b] = pp.breaks[(size_t)b];
                auto put = [&](int f, int q, LD v) {
                    if (q < 8) tp_c8[((size_t)i * nf + f) * 8 + q] = (double)v;
                    else tp_c9[(size_t)i * nf + f] = (double)v;
                };
                for (int q = 0; q < 9; q++) {
                    a0tp[q] += w * pp.coefs[(size_t)0 * 9 + q];  // constant term of the first piece
                    for (int mm = 1; mm <= dj; mm++) put(mm - 1, q, w * pp.coefs[((size_t)mm) * 9 + q]);
                    for (int b = 0; b + 1 < P; b++) {
                        if (spline_only) {
                            put(D + b, q, w * jumps[(size_t)i].d[((size_t)b * (D + 1) + D) * 9 + q]);
                        } else {
                            for (int mm = 0; mm <= D; mm++)
                                put(D + b * (D + 1) + mm, q, w * jumps[(size_t)i].d[((size_t)b * (D + 1) + mm) * 9 + q]);
                        }
                    }
                }
            }
            for (int q = 0; q < 9; q++) m->alpha0_tp[q] = (double)a0tp[q];
        }
    }

    int rc = RN_OK;
    if ((rc = upload(&m->d_ref_wrapped, ref)) != RN_OK || (rc = upload(&m->d_zero_ref, zero_ref)) != RN_OK ||
        (rc = upload(&m->d_g_frac, g_frac_d)) != RN_OK || (rc = upload(&m->d_g_cart, g_cart_d)) != RN_OK ||
        (rc = upload(&m->d_v_frac, v_frac)) != RN_OK || (rc = upload(&m->d_v_cart, v_cart)) != RN_OK ||
        (rc = upload(&m->d_piece_off, piece_off)) != RN_OK || (rc = upload(&m->d_breaks, breaks)) != RN_OK ||
        (rc = upload(&m->d_pieces, pieces)) != RN_OK || (rc = upload(&m->d_tp_x0, tp_x0)) != RN_OK ||
        (rc = upload(&m->d_tp_brk, tp_brk)) != RN_OK || (rc = upload(&m->d_tp_c8, tp_c8)) != RN_OK ||
        (rc = upload(&m->d_tp_c9, tp_c9)) != RN_OK) {
        destroy_model(m);
        return rc;
    }
    *out = m;
    return RN_OK;
}

extern "C" int rn_model_destroy(rn_model* model) {
    if (!model) return RN_OK;
    DeviceGuard guard(model->device);
    destroy_model(model);
    return RN_OK;
}

extern "C" int rn_model_info(const rn_model* model, int64_t info[8]) {
    RN_CHECK_ARG(model && info, "null pointer");
    info[0] = model->num_atoms;
    info[1] = model->num_dofs;
    info[2] = model->num_linear;
    info[3] = model->num_dense;
    info[4] = model->dense_degree;
    info[5] = model->device;
    info[6] = model->affine_kp > 0 ? 1 : 0;
    info[7] = model->dense_max_pieces;
    return RN_OK;
}
