// recip.cu -- reciprocal-space Ewald on the device.
//
// Reference: ReciprocalEwaldState / ReciprocalEwaldGaussian, /root/reference/include/coulombgalore/ewald.hpp:44-94, 214-385:
//   k-vectors    all integer triples n with 0 < |n|^2 <= kc^2, k = 2 pi n / L            (generateKVectors :257-286)
//   A_k          exp(-((|k|^2 + kappa^2) Rc^2 + kappa^2 Rc^2) / (4 alpha^2 Rc^2)) / (|k|^2 + kappa^2)   (setAks :216-225)
//   Q_k          sum_p (z_p + i mu_p.k) exp(i k.r_p)                                      (calcQ :236-250)
//   energy       2 pi / V sum_k |Q_k|^2 A_k                                               (:305-317)
//   force on p   -4 pi / V sum_k Re[exp(i k.r_p) (-mu_p.k + i z_p) conj(Q_k)] k A_k       (:325-336)
//   field at r   -4 pi / V sum_k Re[(-sin k.r + i cos k.r) conj(Q_k)] k A_k               (:342-351)
//   surface      total dipole M = sum z r + sum mu; energy 2 pi |M|^2 / ((2 eps + 1) V), field -4 pi M / ((2 eps + 1) V)
//
// The work is O(N K) complex multiply-adds (K = 5574 for kc = 11): an FP64-pipe kernel with no neighbour structure.
// exp(i k.r) is never a sincos per (particle, k): it is the product of three per-axis phase factors that advance by one
// complex multiplication per step (4 FP64 instructions instead of ~60).
//   * Q: one thread per k-vector; particles stream through shared memory in tiles together with their per-axis tables
//     exp(i m theta), m = 0..kc (built once per tile, one sincos per particle and axis); partial sums per particle slab,
//     then a fixed-order reduction -> bit-reproducible.
//   * force / field: one thread per particle; the k-loop is regenerated from the reference's order (nx, ny outer; nz inner, walked
//     from 0 outwards so that +nz and -nz share one recurrence); Q_k and A_k are warp-uniform loads.
//   * k and -k: Q_{-k} = conj(Q_k) and the force / field terms of k and -k are equal, so both kernels do half of the list.
// Multi-GPU: every rank updates Q with its own particles, the ranks all-reduce the 2K doubles (cg_recip_q_device), then each
// rank evaluates its own particles.  No spatial structure is needed.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "cg_b200.h"

namespace {

constexpr int kQThreads = 256;   // k-vectors per CTA in the Q kernel
constexpr int kTile = 32;        // particles per shared-memory tile
constexpr int kMaxKc = 24;      // tile tables: kTile * 3 * (kc + 1) complex doubles must fit the default 48 KB of shared memory

struct RecipDev {
    int K, kc, lattice;          // lattice = 0: the degenerate single dummy vector of kc = 0 (direct sincos)
    const int *n3;               // [K] packed (nx + 64) | (ny + 64) << 8 | (nz + 64) << 16
    const double *kvec;          // [3K]
    const double *aks;           // [K]
    double two_pi_inv_box[3];
    double prefac;               // -4 pi / V
};

template <class T> struct Buf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---- Q_k = sum_p (z_p + i mu_p.k) exp(i k.r_p) over one slab of particles; grid (ceil(K / 256), nslabs) ----
__global__ void __launch_bounds__(kQThreads) recip_q_kernel(const __grid_constant__ RecipDev rd, long long n, const double *x,
                                                            const double *y, const double *z, const double *q, const double *mx,
                                                            const double *my, const double *mz, double2 *partial) {
    extern __shared__ double2 smem[];
    const int nt = rd.kc + 1;
    double2 *tab = smem;                                                // [kTile][3][nt]
    double *pq = reinterpret_cast<double *>(smem + kTile * 3 * nt);     // [kTile][4]: z, mux, muy, muz
    // Q_{-k} = conj(Q_k) and -k sits at index K - 1 - idx (the list is in lexicographic order): half the list is computed
    const int kk = blockIdx.x * kQThreads + threadIdx.x;
    const bool live = kk < (rd.lattice ? rd.K / 2 : rd.K);
    int nx = 0, ny = 0, nz = 0;
    double kx = 0, ky = 0, kz = 0;
    if (live) {
        const int pk = rd.n3[kk];
        nx = (pk & 255) - 64;
        ny = ((pk >> 8) & 255) - 64;
        nz = ((pk >> 16) & 255) - 64;
        kx = rd.kvec[3 * kk];
        ky = rd.kvec[3 * kk + 1];
        kz = rd.kvec[3 * kk + 2];
    }
    const int ax = abs(nx), ay = abs(ny), az = abs(nz);
    const double sx = nx < 0 ? -1.0 : 1.0, sy = ny < 0 ? -1.0 : 1.0, sz = nz < 0 ? -1.0 : 1.0; // conj for negative indices
    const long long p0 = n * blockIdx.y / gridDim.y, p1 = n * (blockIdx.y + 1) / gridDim.y;
    double qr = 0.0, qi = 0.0;
    for (long long base = p0; base < p1; base += kTile) {
        const int np = (int)min((long long)kTile, p1 - base);
        __syncthreads();
        if (threadIdx.x < 3 * kTile) { // per-axis tables exp(i m theta), m = 0..kc, by recurrence from one sincos
            const int p = threadIdx.x / 3, a = threadIdx.x - 3 * p;
            if (p < np) {
                const double c = a == 0 ? x[base + p] : (a == 1 ? y[base + p] : z[base + p]);
                double sn, cs;
                sincos(rd.two_pi_inv_box[a] * c, &sn, &cs);
                const double2 b = make_double2(cs, sn);
                double2 w = make_double2(1.0, 0.0);
                double2 *t = tab + (p * 3 + a) * nt;
                for (int m = 0; m < nt; m++) {
                    t[m] = w;
                    w = cmul(w, b);
                }
            }
        } else if (threadIdx.x < 4 * kTile) {
            const int p = threadIdx.x - 3 * kTile;
            if (p < np) {
                pq[4 * p + 0] = q ? q[base + p] : 0.0;
                pq[4 * p + 1] = mx ? mx[base + p] : 0.0;
                pq[4 * p + 2] = mx ? my[base + p] : 0.0;
                pq[4 * p + 3] = mx ? mz[base + p] : 0.0;
            }
        }
        __syncthreads();
        if (live) {
            for (int p = 0; p < np; p++) {
                double cs, sn;
                if (rd.lattice) {
                    const double2 *t = tab + p * 3 * nt;
                    double2 a = t[ax], b = t[nt + ay], c = t[2 * nt + az];
                    a.y *= sx;
                    b.y *= sy;
                    c.y *= sz;
                    const double2 e = cmul(cmul(a, b), c);
                    cs = e.x;
                    sn = e.y;
                } else { // the dummy vector of an empty k-list is not a lattice vector
                    sincos(kx * x[base + p] + ky * y[base + p] + kz * z[base + p], &sn, &cs);
                }
                const double zc = pq[4 * p], mk = pq[4 * p + 1] * kx + pq[4 * p + 2] * ky + pq[4 * p + 3] * kz;
                qr += zc * cs - mk * sn;
                qi += zc * sn + mk * cs;
            }
        }
    }
    if (live) partial[(size_t)blockIdx.y * rd.K + kk] = make_double2(qr, qi);
}

// Q = (sign == 0 ? 0 : Q) +- sum over slabs, in slab order; with `mirror` also Q[K - 1 - k] from the conjugate
__global__ void recip_q_reduce_kernel(int K, int nslabs, const double2 *partial, double2 *Q, int sign, int mirror) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= (mirror ? K / 2 : K)) return;
    double sr = 0.0, si = 0.0;
    for (int s = 0; s < nslabs; s++) {
        const double2 v = partial[(size_t)s * K + kk];
        sr += v.x;
        si += v.y;
    }
    const double sg = sign < 0 ? -1.0 : 1.0;
    double2 o = sign == 0 ? make_double2(0.0, 0.0) : Q[kk];
    o.x += sg * sr;
    o.y += sg * si;
    Q[kk] = o;
    if (mirror) {
        double2 m = sign == 0 ? make_double2(0.0, 0.0) : Q[K - 1 - kk];
        m.x += sg * sr;
        m.y -= sg * si;
        Q[K - 1 - kk] = m;
    }
}

// energy = 2 pi / V sum_k |Q_k|^2 A_k: one CTA, fixed order
__global__ void recip_energy_kernel(int K, const double2 *Q, const double *aks, double scale, double *out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int k = threadIdx.x; k < K; k += 1024) s += (Q[k].x * Q[k].x + Q[k].y * Q[k].y) * aks[k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = scale * sh[0];
}

// ---- force and field of every particle; one thread per particle, k-loop in the reference's order ----
// row[(nx + kc) * (2 kc + 1) + (ny + kc)] = {first k index of the (nx, ny) column, nzmax or -1 when the column is empty}
template <bool FORCE, bool FIELD>
__global__ void __launch_bounds__(128) recip_force_kernel(const __grid_constant__ RecipDev rd, const int2 *row, const double2 *Q,
                                                          long long n, const double *x, const double *y, const double *z,
                                                          const double *q, const double *mx, const double *my, const double *mz,
                                                          double *force, double *field) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int kc = rd.kc, side = 2 * kc + 1;
    const double zc = q ? q[p] : 0.0;
    const double m0 = mx ? mx[p] : 0.0, m1 = mx ? my[p] : 0.0, m2 = mx ? mz[p] : 0.0;
    double2 bx, by, bz, ex, ey0;
    sincos(rd.two_pi_inv_box[0] * x[p], &bx.y, &bx.x);
    sincos(rd.two_pi_inv_box[1] * y[p], &by.y, &by.x);
    sincos(rd.two_pi_inv_box[2] * z[p], &bz.y, &bz.x);
    sincos(-(double)kc * rd.two_pi_inv_box[1] * y[p], &ey0.y, &ey0.x); // exp(i ny theta_y) at ny = -kc
    double Fx = 0, Fy = 0, Fz = 0, Ex = 0, Ey = 0, Ez = 0;
    // The terms of k and -k are equal (conjugate phase, conjugate Q, opposite k), so only the half space nx > 0, or nx = 0 and
    // ny > 0, or nx = ny = 0 and nz > 0 is walked and the sums are doubled.
    ex = make_double2(1.0, 0.0);
    for (int nx = 0; nx <= kc; nx++, ex = cmul(ex, bx)) {
        const double kx = rd.two_pi_inv_box[0] * nx;
        double2 ey = nx == 0 ? make_double2(1.0, 0.0) : ey0;
        for (int ny = nx == 0 ? 0 : -kc; ny <= kc; ny++, ey = cmul(ey, by)) {
            const int2 r = row[(nx + kc) * side + (ny + kc)];
            if (r.y < 0) continue;
            const double ky = rd.two_pi_inv_box[1] * ny;
            const double2 exy = cmul(ex, ey);
            const double mxy = m0 * kx + m1 * ky;
            const bool origin_col = nx == 0 && ny == 0; // the column that has no nz = 0 and whose nz < 0 half is mirrored
            // k index of (nx, ny, nz): r.x + nz + nzmax, minus one past the skipped origin
            double tf = 0, te = 0, tfz = 0, tez = 0;
            double2 w = make_double2(1.0, 0.0);
            for (int m = 0; m <= r.y; m++, w = cmul(w, bz)) {
#pragma unroll
                for (int sgn = 1; sgn >= -1; sgn -= 2) {
                    if ((m == 0 || origin_col) && (sgn < 0 || (m == 0 && origin_col))) continue;
                    const int nz = sgn * m;
                    const int idx = r.x + nz + r.y - ((origin_col && nz > 0) ? 1 : 0);
                    const double2 e = cmul(exy, make_double2(w.x, sgn * w.y));
                    const double2 Qk = Q[idx];
                    const double ak = rd.aks[idx];
                    const double kz = rd.two_pi_inv_box[2] * nz;
                    if (FORCE) {
                        const double a = -(mxy + m2 * kz); // qmu = (-mu.k, z)
                        const double t = ((e.x * a - e.y * zc) * Qk.x + (e.x * zc + e.y * a) * Qk.y) * ak;
                        tf += t;
                        tfz += t * kz;
                    }
                    if (FIELD) {
                        const double t = (-e.y * Qk.x + e.x * Qk.y) * ak;
                        te += t;
                        tez += t * kz;
                    }
                }
            }
            if (FORCE) {
                Fx += tf * kx;
                Fy += tf * ky;
                Fz += tfz;
            }
            if (FIELD) {
                Ex += te * kx;
                Ey += te * ky;
                Ez += tez;
            }
        }
    }
    Fx *= 2.0; Fy *= 2.0; Fz *= 2.0;
    Ex *= 2.0; Ey *= 2.0; Ez *= 2.0;
    if (FORCE) {
        force[3 * p + 0] = rd.prefac * Fx;
        force[3 * p + 1] = rd.prefac * Fy;
        force[3 * p + 2] = rd.prefac * Fz;
    }
    if (FIELD) {
        field[3 * p + 0] = rd.prefac * Ex;
        field[3 * p + 1] = rd.prefac * Ey;
        field[3 * p + 2] = rd.prefac * Ez;
    }
}

// the degenerate k-list (kc = 0: one dummy vector with zero weight): direct evaluation
__global__ void recip_force_generic_kernel(const __grid_constant__ RecipDev rd, const double2 *Q, long long n, const double *x,
                                           const double *y, const double *z, const double *q, const double *mx, const double *my,
                                           const double *mz, double *force, double *field) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double F[3] = {0, 0, 0}, E[3] = {0, 0, 0};
    for (int k = 0; k < rd.K; k++) {
        const double kx = rd.kvec[3 * k], ky = rd.kvec[3 * k + 1], kz = rd.kvec[3 * k + 2];
        double sn, cs;
        sincos(kx * x[p] + ky * y[p] + kz * z[p], &sn, &cs);
        const double a = -(mx ? mx[p] * kx + my[p] * ky + mz[p] * kz : 0.0), zc = q ? q[p] : 0.0;
        const double tf = ((cs * a - sn * zc) * Q[k].x + (cs * zc + sn * a) * Q[k].y) * rd.aks[k];
        const double te = (-sn * Q[k].x + cs * Q[k].y) * rd.aks[k];
        F[0] += tf * kx; F[1] += tf * ky; F[2] += tf * kz;
        E[0] += te * kx; E[1] += te * ky; E[2] += te * kz;
    }
    for (int d = 0; d < 3; d++) {
        if (force) force[3 * p + d] = rd.prefac * F[d];
        if (field) field[3 * p + d] = rd.prefac * E[d];
    }
}

// total dipole M = sum z r + sum mu: per-block partials, then one thread in block order
__global__ void recip_dipole_partial_kernel(long long n, const double *x, const double *y, const double *z, const double *q,
                                            const double *mx, const double *my, const double *mz, double *partial) {
    __shared__ double sh[3][256];
    double a = 0, b = 0, c = 0;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n; p += (long long)gridDim.x * 256) {
        const double zc = q ? q[p] : 0.0;
        a += x[p] * zc + (mx ? mx[p] : 0.0);
        b += y[p] * zc + (mx ? my[p] : 0.0);
        c += z[p] * zc + (mx ? mz[p] : 0.0);
    }
    sh[0][threadIdx.x] = a;
    sh[1][threadIdx.x] = b;
    sh[2][threadIdx.x] = c;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off)
            for (int d = 0; d < 3; d++) sh[d][threadIdx.x] += sh[d][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int d = 0; d < 3; d++) partial[3 * blockIdx.x + d] = sh[d][0];
}
__global__ void recip_dipole_final_kernel(int nb, const double *partial, double *out3) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double M[3] = {0, 0, 0};
    for (int b = 0; b < nb; b++)
        for (int d = 0; d < 3; d++) M[d] += partial[3 * b + d];
    for (int d = 0; d < 3; d++) out3[d] = M[d];
}

} // namespace

struct cg_recip {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    double cutoff = 0, alpha = 0, eps_sur = 0, kappa = 0, box[3] = {0, 0, 0};
    int kc = 0, K = 0, lattice = 1;
    std::vector<double> h_k, h_aks;
    Buf<int> d_n3;
    Buf<int2> d_row;
    Buf<double> d_k, d_aks, d_scalars, d_dip_partial, d_out_force, d_out_field;
    Buf<double2> d_q, d_partial;
    Buf<double> d_in[7];
    double *h_pinned = nullptr;
    std::string err;
};

static thread_local std::string g_recip_error;

#define RC_CUDA(r, call)                                                                                            \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess) {                                                                                   \
            (r)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                         \
            return (e__ == cudaErrorMemoryAllocation) ? CG_ERR_ALLOC : CG_ERR_CUDA;                                 \
        }                                                                                                           \
    } while (0)

static RecipDev dev_params(const cg_recip *r) {
    RecipDev rd;
    const double pi = 4.0 * std::atan(1.0);
    rd.K = r->K;
    rd.kc = r->kc;
    rd.lattice = r->lattice;
    rd.n3 = r->d_n3.p;
    rd.kvec = r->d_k.p;
    rd.aks = r->d_aks.p;
    for (int d = 0; d < 3; d++) rd.two_pi_inv_box[d] = 2.0 * pi * (1.0 / r->box[d]);
    rd.prefac = -4.0 * pi / (r->box[0] * r->box[1] * r->box[2]);
    return rd;
}

extern "C" {

const char *cg_recip_last_error(const cg_recip *r) { return r ? r->err.c_str() : g_recip_error.c_str(); }

int cg_recip_create(double cutoff, double alpha, double eps_sur, double kappa, int reciprocal_cutoff, const double *box, int device,
                    cg_recip **out) {
    if (!out || !box || !(cutoff > 0.0) || !(alpha > 0.0) || reciprocal_cutoff < 0 || reciprocal_cutoff > kMaxKc || !(box[0] > 0.0) ||
        !(box[1] > 0.0) || !(box[2] > 0.0))
        return CG_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_recip_error = "no CUDA device: this library has no CPU fallback";
        return CG_ERR_NO_DEVICE;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return CG_ERR_NO_DEVICE;
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return CG_ERR_INVALID;
    cg_recip *r = new cg_recip();
    r->device = device;
    r->cutoff = cutoff;
    r->alpha = alpha;
    r->eps_sur = eps_sur;
    r->kappa = kappa;
    r->kc = reciprocal_cutoff;
    for (int d = 0; d < 3; d++) r->box[d] = box[d];
    const double pi = 4.0 * std::atan(1.0);
    const int kc = r->kc, side = 2 * kc + 1;
    std::vector<int> n3;
    std::vector<int2> row((size_t)side * side, make_int2(0, -1));
    // generateKVectors (:257-286): nx outermost, nz innermost; every (nx, ny) column is a contiguous range of nz
    for (int nx = -kc; nx <= kc; nx++)
        for (int ny = -kc; ny <= kc; ny++) {
            int2 &rw = row[(size_t)(nx + kc) * side + (ny + kc)];
            for (int nz = -kc; nz <= kc; nz++) {
                const int rr = nx * nx + ny * ny + nz * nz;
                if (rr > 0 && rr <= kc * kc) {
                    if (rw.y < 0) rw = make_int2((int)n3.size(), 0);
                    rw.y = std::max(rw.y, std::abs(nz));
                    n3.push_back((nx + 64) | ((ny + 64) << 8) | ((nz + 64) << 16));
                    r->h_k.push_back(2.0 * pi * (1.0 / box[0]) * nx);
                    r->h_k.push_back(2.0 * pi * (1.0 / box[1]) * ny);
                    r->h_k.push_back(2.0 * pi * (1.0 / box[2]) * nz);
                }
            }
        }
    if (n3.empty()) { // resize(0) (:68-72): one dummy vector (1, 0, 0) with zero weight
        r->lattice = 0;
        n3.push_back(64 | (64 << 8) | (64 << 16));
        r->h_k = {1.0, 0.0, 0.0};
        r->h_aks = {0.0};
    } else { // setAks (:216-225); kappa^2 enters twice, as in the reference
        const double c2 = cutoff * cutoff, rd2 = alpha * alpha * c2, rk2 = kappa * kappa * c2;
        r->h_aks.resize(n3.size());
        for (size_t i = 0; i < n3.size(); i++) {
            const double *k = &r->h_k[3 * i];
            const double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2] + kappa * kappa;
            r->h_aks[i] = std::exp(-(k2 * c2 + rk2) / (4.0 * rd2)) / k2;
        }
    }
    r->K = (int)n3.size();
    bool ok = cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    r->stream = r->own_stream;
    ok = ok && r->d_n3.ensure(n3.size()) == cudaSuccess && r->d_row.ensure(row.size()) == cudaSuccess &&
         r->d_k.ensure(r->h_k.size()) == cudaSuccess && r->d_aks.ensure(r->h_aks.size()) == cudaSuccess &&
         r->d_q.ensure((size_t)r->K) == cudaSuccess && r->d_scalars.ensure(8) == cudaSuccess &&
         cudaMallocHost((void **)&r->h_pinned, 8 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMemcpy(r->d_n3.p, n3.data(), n3.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(r->d_row.p, row.data(), row.size() * sizeof(int2), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(r->d_k.p, r->h_k.data(), r->h_k.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(r->d_aks.p, r->h_aks.data(), r->h_aks.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemset(r->d_q.p, 0, (size_t)r->K * sizeof(double2)) == cudaSuccess;
    if (!ok) {
        g_recip_error = std::string("cg_recip_create: ") + cudaGetErrorString(cudaGetLastError());
        cg_recip_destroy(r);
        return CG_ERR_CUDA;
    }
    *out = r;
    return CG_OK;
}

int cg_recip_destroy(cg_recip *r) {
    if (!r) return CG_OK;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    r->d_n3.release();
    r->d_row.release();
    r->d_k.release();
    r->d_aks.release();
    r->d_scalars.release();
    r->d_dip_partial.release();
    r->d_out_force.release();
    r->d_out_field.release();
    r->d_q.release();
    r->d_partial.release();
    for (auto &b : r->d_in) b.release();
    if (r->h_pinned) cudaFreeHost(r->h_pinned);
    if (r->own_stream) cudaStreamDestroy(r->own_stream);
    delete r;
    return CG_OK;
}

int cg_recip_set_stream(cg_recip *r, void *cuda_stream) {
    if (!r) return CG_ERR_INVALID;
    r->stream = cuda_stream ? (cudaStream_t)cuda_stream : r->own_stream;
    return CG_OK;
}

int cg_recip_num_kvectors(const cg_recip *r) { return r ? r->K : 0; }

int cg_recip_kvectors(const cg_recip *r, double *k3, double *aks) {
    if (!r) return CG_ERR_INVALID;
    if (k3) std::copy(r->h_k.begin(), r->h_k.end(), k3);
    if (aks) std::copy(r->h_aks.begin(), r->h_aks.end(), aks);
    return CG_OK;
}

double *cg_recip_q_device(cg_recip *r) { return r ? reinterpret_cast<double *>(r->d_q.p) : nullptr; }

int cg_recip_update_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                           const double *mux, const double *muy, const double *muz, int sign) {
    if (!r || n < 0 || (n > 0 && (!x || !y || !z))) return CG_ERR_INVALID;
    if ((mux || muy || muz) && !(mux && muy && muz)) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const RecipDev rd = dev_params(r);
    const int kwork = r->lattice ? r->K / 2 : r->K;
    const int kblocks = (kwork + kQThreads - 1) / kQThreads;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device);
    // particle slabs: enough CTAs for ~4 per SM, never fewer than one tile of particles per slab
    int nslabs = (int)std::max<int64_t>(1, std::min<int64_t>((4 * sms + kblocks - 1) / kblocks, (n + kTile - 1) / kTile));
    RC_CUDA(r, r->d_partial.ensure((size_t)nslabs * r->K));
    const size_t smem = (size_t)kTile * 3 * (r->kc + 1) * sizeof(double2) + (size_t)kTile * 4 * sizeof(double);
    recip_q_kernel<<<dim3(kblocks, nslabs), kQThreads, smem, r->stream>>>(rd, n, x, y, z, q, mux, muy, muz, r->d_partial.p);
    recip_q_reduce_kernel<<<(kwork + 255) / 256, 256, 0, r->stream>>>(r->K, nslabs, r->d_partial.p, r->d_q.p, sign, r->lattice);
    RC_CUDA(r, cudaGetLastError());
    return CG_OK;
}

int cg_recip_energy_device(cg_recip *r, double *energy_dev) {
    if (!r || !energy_dev) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const double pi = 4.0 * std::atan(1.0);
    recip_energy_kernel<<<1, 1024, 0, r->stream>>>(r->K, r->d_q.p, r->d_aks.p, 2.0 * pi / (r->box[0] * r->box[1] * r->box[2]), energy_dev);
    RC_CUDA(r, cudaGetLastError());
    return CG_OK;
}

int cg_recip_force_field_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                                const double *mux, const double *muy, const double *muz, double *force, double *field) {
    if (!r || n < 0 || (n > 0 && (!x || !y || !z))) return CG_ERR_INVALID;
    if ((mux || muy || muz) && !(mux && muy && muz)) return CG_ERR_INVALID;
    if (n == 0 || (!force && !field)) return CG_OK;
    cudaSetDevice(r->device);
    const RecipDev rd = dev_params(r);
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (!r->lattice)
        recip_force_generic_kernel<<<grid, 128, 0, r->stream>>>(rd, r->d_q.p, n, x, y, z, q, mux, muy, muz, force, field);
    else if (force && field)
        recip_force_kernel<true, true><<<grid, 128, 0, r->stream>>>(rd, r->d_row.p, r->d_q.p, n, x, y, z, q, mux, muy, muz, force, field);
    else if (force)
        recip_force_kernel<true, false><<<grid, 128, 0, r->stream>>>(rd, r->d_row.p, r->d_q.p, n, x, y, z, q, mux, muy, muz, force, field);
    else
        recip_force_kernel<false, true><<<grid, 128, 0, r->stream>>>(rd, r->d_row.p, r->d_q.p, n, x, y, z, q, mux, muy, muz, force, field);
    RC_CUDA(r, cudaGetLastError());
    return CG_OK;
}

// dipole3_dev[0..2] = total dipole M = sum z r + sum mu (dipoleMoment :77-88); the surface terms follow on the host
int cg_recip_surface_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux,
                            const double *muy, const double *muz, double *dipole3_dev) {
    if (!r || !dipole3_dev || n < 0 || (n > 0 && (!x || !y || !z))) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(512, (n + 255) / 256));
    RC_CUDA(r, r->d_dip_partial.ensure(3 * (size_t)nb));
    recip_dipole_partial_kernel<<<nb, 256, 0, r->stream>>>(n, x, y, z, q, mux, muy, muz, r->d_dip_partial.p);
    recip_dipole_final_kernel<<<1, 32, 0, r->stream>>>(nb, r->d_dip_partial.p, dipole3_dev);
    RC_CUDA(r, cudaGetLastError());
    return CG_OK;
}

// ---- host-buffer forms ----
static int upload(cg_recip *r, int64_t n, const double *const src[7], const double *dst[7]) {
    for (int k = 0; k < 7; k++) {
        dst[k] = nullptr;
        if (src[k] && n > 0) {
            RC_CUDA(r, r->d_in[k].ensure((size_t)n));
            RC_CUDA(r, cudaMemcpyAsync(r->d_in[k].p, src[k], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, r->stream));
            dst[k] = r->d_in[k].p;
        }
    }
    return CG_OK;
}

int cg_recip_update(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux,
                    const double *muy, const double *muz, int sign) {
    if (!r) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const double *src[7] = {x, y, z, q, mux, muy, muz}, *d[7];
    int rc = upload(r, n, src, d);
    if (rc != CG_OK) return rc;
    rc = cg_recip_update_device(r, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], sign);
    if (rc != CG_OK) return rc;
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    return CG_OK;
}

int cg_recip_get_q(cg_recip *r, double *q2k) {
    if (!r || !q2k) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    RC_CUDA(r, cudaMemcpyAsync(q2k, r->d_q.p, (size_t)r->K * sizeof(double2), cudaMemcpyDeviceToHost, r->stream));
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    return CG_OK;
}

int cg_recip_set_q(cg_recip *r, const double *q2k) {
    if (!r || !q2k) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    RC_CUDA(r, cudaMemcpyAsync(r->d_q.p, q2k, (size_t)r->K * sizeof(double2), cudaMemcpyHostToDevice, r->stream));
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    return CG_OK;
}

int cg_recip_energy(cg_recip *r, double *energy) {
    if (!r || !energy) return CG_ERR_INVALID;
    const int rc = cg_recip_energy_device(r, r->d_scalars.p);
    if (rc != CG_OK) return rc;
    RC_CUDA(r, cudaMemcpyAsync(r->h_pinned, r->d_scalars.p, sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    *energy = r->h_pinned[0];
    return CG_OK;
}

int cg_recip_force_field(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux,
                         const double *muy, const double *muz, double *force, double *field) {
    if (!r) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const double *src[7] = {x, y, z, q, mux, muy, muz}, *d[7];
    int rc = upload(r, n, src, d);
    if (rc != CG_OK) return rc;
    if (force) RC_CUDA(r, r->d_out_force.ensure(3 * (size_t)n + 1));
    if (field) RC_CUDA(r, r->d_out_field.ensure(3 * (size_t)n + 1));
    rc = cg_recip_force_field_device(r, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], force ? r->d_out_force.p : nullptr,
                                     field ? r->d_out_field.p : nullptr);
    if (rc != CG_OK) return rc;
    if (force && n > 0) RC_CUDA(r, cudaMemcpyAsync(force, r->d_out_force.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    if (field && n > 0) RC_CUDA(r, cudaMemcpyAsync(field, r->d_out_field.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    return CG_OK;
}

// surface_energy (:367-374) and surface_field (:355-362) of the particle set
int cg_recip_surface(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux,
                     const double *muy, const double *muz, double *energy, double *field3) {
    if (!r) return CG_ERR_INVALID;
    cudaSetDevice(r->device);
    const double *src[7] = {x, y, z, q, mux, muy, muz}, *d[7];
    int rc = upload(r, n, src, d);
    if (rc != CG_OK) return rc;
    rc = cg_recip_surface_device(r, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], r->d_scalars.p);
    if (rc != CG_OK) return rc;
    RC_CUDA(r, cudaMemcpyAsync(r->h_pinned, r->d_scalars.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    RC_CUDA(r, cudaStreamSynchronize(r->stream));
    const double pi = 4.0 * std::atan(1.0), V = r->box[0] * r->box[1] * r->box[2];
    const double *M = r->h_pinned;
    if (energy) *energy = 2.0 * pi / (2.0 * r->eps_sur + 1.0) / V * (M[0] * M[0] + M[1] * M[1] + M[2] * M[2]);
    if (field3)
        for (int k = 0; k < 3; k++) field3[k] = -4.0 * pi / (2.0 * r->eps_sur + 1.0) / V * M[k];
    return CG_OK;
}

} // extern "C"
