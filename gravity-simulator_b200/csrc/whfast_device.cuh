// The WHFast interaction acceleration of ONE target particle (device function shared by whfast.cu's kernel and the
// fused tail kernel of whfast_resident.cu).  See whfast.cu for the reference citations.
#pragma once
#include "internal.cuh"

namespace gb {

struct V3 { double x, y, z; };

__device__ __forceinline__ double norm3(double x, double y, double z)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}
// d = x[to] - x[from]; returns |d|^3 + eps^3
// (posm is deliberately not __restrict__: the fused tail kernel of whfast_resident.cu writes the target's own record
// just before calling this, and a read-only-cache load must not be allowed to miss that store)
__device__ __forceinline__ double sep_cubed(V3 &d, const double4 *posm, int to, int from, double eps3)
{
    const double4 a = posm[to], b = posm[from];
    d.x = __dsub_rn(a.x, b.x); d.y = __dsub_rn(a.y, b.y); d.z = __dsub_rn(a.z, b.z);
    const double r = norm3(d.x, d.y, d.z);
    return __dadd_rn(__dmul_rn(__dmul_rn(r, r), r), eps3);
}
// s += (gm * d) / c   with gm already = ((G*m_j)[*m_k])
__device__ __forceinline__ void add_term(V3 &s, double gm, const V3 &d, double c)
{
    s.x = __dadd_rn(s.x, __ddiv_rn(__dmul_rn(gm, d.x), c));
    s.y = __dadd_rn(s.y, __ddiv_rn(__dmul_rn(gm, d.y), c));
    s.z = __dadd_rn(s.z, __ddiv_rn(__dmul_rn(gm, d.z), c));
}

// sum over massive pairs (j, k) with j < i < k of G m_j m_k (x_k - x_j) / (|x_k - x_j|^3 + eps^3), in the reference's loop
// order (:1225-1252); i is a massless target
__device__ __forceinline__ void whfast_massless_pair_sum(V3 &s, int i, const double4 *posm, double G, double eps3,
                                                         const int *__restrict__ list, int nl)
{
    V3 d;
    for (int q = 0; q < nl; q++) {
        const int j = list[q];
        if (j >= i) break;
        const double gmj = __dmul_rn(G, posm[j].w);
        for (int r = q + 1; r < nl; r++) {
            const int k = list[r];
            if (k <= i) continue;
            const double c = sep_cubed(d, posm, k, j, eps3);
            add_term(s, __dmul_rn(gmj, posm[k].w), d, c);
        }
    }
}

// list == nullptr: "pairwise" variant, the particle list is 0..n-1.  Otherwise list[0..nl) are the massive
// particle ids in index order and rank[i] is the list position of massive particle i.
__device__ __forceinline__ void whfast_accel_one(int i, int n, const double4 *posm, double G,
                                                 const double *__restrict__ jx, const double *__restrict__ eta,
                                                 double eps3, const int *__restrict__ list, int nl,
                                                 const int *__restrict__ rank, double *acc,
                                                 const double *__restrict__ pair_sum = nullptr)
{
    const bool massless_variant = list != nullptr;
    const double mi = posm[i].w;
    const bool target_massive = !massless_variant || mi != 0.0;
    int p = i;                       // position of the target in the list (massive targets only)
    if (massless_variant && target_massive) p = rank[i];
    if (target_massive ? (p == 0) : (i == 0)) return;   // the reference's loops start at 1 / skip particle 0

#define LIST(q) (massless_variant ? list[q] : (q))
    const int L = massless_variant ? nl : n;
    const double m0 = posm[0].w;
    V3 d;
    // central body + Jacobi term, written straight into a[i] (:866-883, :1033-1050, :1152-1169)
    const double c0 = sep_cubed(d, posm, i, 0, eps3);
    const double jxx = jx[3 * (size_t)i], jxy = jx[3 * (size_t)i + 1], jxz = jx[3 * (size_t)i + 2];
    const double jn = norm3(jxx, jxy, jxz);
    const double cj = __dadd_rn(__dmul_rn(__dmul_rn(jn, jn), jn), eps3);
    const double eta_i = eta[i], eta_im1 = eta[i - 1];
    double pre = __dmul_rn(G, m0);
    if (target_massive) pre = __ddiv_rn(__dmul_rn(pre, eta_i), eta_im1);
    V3 a;
    a.x = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxx, cj), __ddiv_rn(d.x, c0)));
    a.y = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxy, cj), __ddiv_rn(d.y, c0)));
    a.z = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxz, cj), __ddiv_rn(d.z, c0)));

    V3 s = {0.0, 0.0, 0.0};
    if (target_massive) {
        for (int q = 1; q < p; q++) {                      // bodies inside the target's orbit
            const int j = LIST(q);
            const double c = sep_cubed(d, posm, i, j, eps3);
            add_term(s, __dmul_rn(G, posm[j].w), d, c);
        }
        a.x = __dsub_rn(a.x, __ddiv_rn(__dmul_rn(s.x, eta_i), eta_im1));
        a.y = __dsub_rn(a.y, __ddiv_rn(__dmul_rn(s.y, eta_i), eta_im1));
        a.z = __dsub_rn(a.z, __ddiv_rn(__dmul_rn(s.z, eta_i), eta_im1));
        s.x = s.y = s.z = 0.0;
        for (int q = p + 1; q < L; q++) {                  // bodies outside
            const int j = LIST(q);
            const double c = sep_cubed(d, posm, j, i, eps3);
            add_term(s, __dmul_rn(G, posm[j].w), d, c);
        }
        a.x = __dadd_rn(a.x, s.x); a.y = __dadd_rn(a.y, s.y); a.z = __dadd_rn(a.z, s.z);
        s.x = s.y = s.z = 0.0;
        for (int q = 0; q < p; q++) {                      // pairs straddling the target
            const int j = LIST(q);
            const double gmj = __dmul_rn(G, posm[j].w);
            for (int r = p + 1; r < L; r++) {
                const int k = LIST(r);
                const double c = sep_cubed(d, posm, k, j, eps3);
                add_term(s, __dmul_rn(gmj, posm[k].w), d, c);
            }
        }
    } else {
        for (int q = 1; q < nl; q++) {                     // massive bodies before the target (:1171-1189)
            const int j = list[q];
            if (j >= i) break;
            const double c = sep_cubed(d, posm, i, j, eps3);
            add_term(s, __dmul_rn(G, posm[j].w), d, c);
        }
        a.x = __dsub_rn(a.x, s.x); a.y = __dsub_rn(a.y, s.y); a.z = __dsub_rn(a.z, s.z);
        s.x = s.y = s.z = 0.0;
        for (int q = 1; q < nl; q++) {                     // massive bodies after the target (:1198-1216)
            const int j = list[q];
            if (j <= i) continue;
            const double c = sep_cubed(d, posm, j, i, eps3);
            add_term(s, __dmul_rn(G, posm[j].w), d, c);
        }
        a.x = __dadd_rn(a.x, s.x); a.y = __dadd_rn(a.y, s.y); a.z = __dadd_rn(a.z, s.z);
        s.x = s.y = s.z = 0.0;
        if (pair_sum) {
            // the straddling-pair sum only depends on how many massive particles precede the target: the resident step
            // evaluates it once per gap (whfast_resident.cu, wh_j2c_skel_kernel) with the loop below
            const int g = rank[i];
            s.x = pair_sum[3 * g + 0]; s.y = pair_sum[3 * g + 1]; s.z = pair_sum[3 * g + 2];
        } else {
            whfast_massless_pair_sum(s, i, posm, G, eps3, list, nl);
        }
    }
    a.x = __dsub_rn(a.x, __ddiv_rn(s.x, eta_im1));
    a.y = __dsub_rn(a.y, __ddiv_rn(s.y, eta_im1));
    a.z = __dsub_rn(a.z, __ddiv_rn(s.z, eta_im1));
#undef LIST
    acc[3 * (size_t)i + 0] = a.x;
    acc[3 * (size_t)i + 1] = a.y;
    acc[3 * (size_t)i + 2] = a.z;
}


// ---- warp-cooperative forms (whfast_resident.cu) ------------------------------------------------------------------
// A massive target's three sums are chains of 10-40 terms, each a square root and three divisions: one thread spends
// ~0.6 us per term on FP64 latency alone.  The terms are independent; only the additions are ordered.  So the 32 lanes
// evaluate 32 terms at a time and then every lane adds them up in the reference's order (bit-identical result).
template <class F>
__device__ __forceinline__ V3 warp_ordered_sum(int nterms, F term)
{
    const int lane = threadIdx.x & 31;
    V3 s = {0.0, 0.0, 0.0};
    for (int t0 = 0; t0 < nterms; t0 += 32) {
        V3 v = {0.0, 0.0, 0.0};
        if (t0 + lane < nterms) v = term(t0 + lane);
        const int cnt = min(32, nterms - t0);
        for (int k = 0; k < cnt; k++) {
            s.x = __dadd_rn(s.x, __shfl_sync(0xffffffffu, v.x, k));
            s.y = __dadd_rn(s.y, __shfl_sync(0xffffffffu, v.y, k));
            s.z = __dadd_rn(s.z, __shfl_sync(0xffffffffu, v.z, k));
        }
    }
    return s;
}
// (gm * d) / c per component: the quantity add_term() adds
__device__ __forceinline__ V3 quotient_term(double gm, const V3 &d, double c)
{
    V3 t;
    t.x = __ddiv_rn(__dmul_rn(gm, d.x), c);
    t.y = __ddiv_rn(__dmul_rn(gm, d.y), c);
    t.z = __ddiv_rn(__dmul_rn(gm, d.z), c);
    return t;
}

// Massless method, MASSIVE target at list position p (called by a whole warp; :1006-1128).  Writes acc[i].
__device__ __forceinline__ void whfast_accel_massive_warp(int p, const double4 *posm, double G, const double *__restrict__ jx,
                                                          const double *__restrict__ eta, double eps3, const int *__restrict__ list,
                                                          int nl, double *acc)
{
    if (p == 0) return;                        // the reference's loop starts at 1
    const int i = list[p];
    const double m0 = posm[0].w;
    V3 d0;
    const double c0 = sep_cubed(d0, posm, i, 0, eps3);
    const double jxx = jx[3 * (size_t)i], jxy = jx[3 * (size_t)i + 1], jxz = jx[3 * (size_t)i + 2];
    const double jn = norm3(jxx, jxy, jxz);
    const double cj = __dadd_rn(__dmul_rn(__dmul_rn(jn, jn), jn), eps3);
    const double eta_i = eta[i], eta_im1 = eta[i - 1];
    const double pre = __ddiv_rn(__dmul_rn(__dmul_rn(G, m0), eta_i), eta_im1);
    V3 a;
    a.x = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxx, cj), __ddiv_rn(d0.x, c0)));
    a.y = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxy, cj), __ddiv_rn(d0.y, c0)));
    a.z = __dmul_rn(pre, __dsub_rn(__ddiv_rn(jxz, cj), __ddiv_rn(d0.z, c0)));
    V3 s = warp_ordered_sum(p - 1, [&](int t) {                    // bodies inside the target's orbit
        const int j = list[1 + t];
        V3 d;
        const double c = sep_cubed(d, posm, i, j, eps3);
        return quotient_term(__dmul_rn(G, posm[j].w), d, c);
    });
    a.x = __dsub_rn(a.x, __ddiv_rn(__dmul_rn(s.x, eta_i), eta_im1));
    a.y = __dsub_rn(a.y, __ddiv_rn(__dmul_rn(s.y, eta_i), eta_im1));
    a.z = __dsub_rn(a.z, __ddiv_rn(__dmul_rn(s.z, eta_i), eta_im1));
    s = warp_ordered_sum(nl - p - 1, [&](int t) {                  // bodies outside
        const int j = list[p + 1 + t];
        V3 d;
        const double c = sep_cubed(d, posm, j, i, eps3);
        return quotient_term(__dmul_rn(G, posm[j].w), d, c);
    });
    a.x = __dadd_rn(a.x, s.x); a.y = __dadd_rn(a.y, s.y); a.z = __dadd_rn(a.z, s.z);
    const int nout = nl - p - 1;
    s = warp_ordered_sum(p * nout, [&](int t) {                    // pairs straddling the target, q-major
        const int q = t / nout, r = p + 1 + t % nout;
        const int j = list[q], k = list[r];
        V3 d;
        const double c = sep_cubed(d, posm, k, j, eps3);
        return quotient_term(__dmul_rn(__dmul_rn(G, posm[j].w), posm[k].w), d, c);
    });
    a.x = __dsub_rn(a.x, __ddiv_rn(s.x, eta_im1));
    a.y = __dsub_rn(a.y, __ddiv_rn(s.y, eta_im1));
    a.z = __dsub_rn(a.z, __ddiv_rn(s.z, eta_im1));
    if ((threadIdx.x & 31) == 0) {
        acc[3 * (size_t)i + 0] = a.x; acc[3 * (size_t)i + 1] = a.y; acc[3 * (size_t)i + 2] = a.z;
    }
}

// Sum over the massive pairs straddling gap g (massive list positions q < g <= r), q-major (:1225-1252), by a warp.
__device__ __forceinline__ V3 whfast_gap_pair_sum_warp(int g, const double4 *posm, double G, double eps3,
                                                       const int *__restrict__ list, int nl)
{
    const int nout = nl - g;
    return warp_ordered_sum(g * nout, [&](int t) {
        const int q = t / nout, r = g + t % nout;
        const int j = list[q], k = list[r];
        V3 d;
        const double c = sep_cubed(d, posm, k, j, eps3);
        return quotient_term(__dmul_rn(__dmul_rn(G, posm[j].w), posm[k].w), d, c);
    });
}

}  // namespace gb
