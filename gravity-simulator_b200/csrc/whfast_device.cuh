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
__device__ __forceinline__ double sep_cubed(V3 &d, const double4 *__restrict__ posm, int to, int from, double eps3)
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

// list == nullptr: "pairwise" variant, the particle list is 0..n-1.  Otherwise list[0..nl) are the massive
// particle ids in index order and rank[i] is the list position of massive particle i.
__device__ __forceinline__ void whfast_accel_one(int i, int n, const double4 *__restrict__ posm, double G,
                                                 const double *__restrict__ jx, const double *__restrict__ eta,
                                                 double eps3, const int *__restrict__ list, int nl,
                                                 const int *__restrict__ rank, double *__restrict__ acc)
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
        for (int q = 0; q < nl; q++) {                     // massive pairs straddling the target (:1225-1252)
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
    a.x = __dsub_rn(a.x, __ddiv_rn(s.x, eta_im1));
    a.y = __dsub_rn(a.y, __ddiv_rn(s.y, eta_im1));
    a.z = __dsub_rn(a.z, __ddiv_rn(s.z, eta_im1));
#undef LIST
    acc[3 * (size_t)i + 0] = a.x;
    acc[3 * (size_t)i + 1] = a.y;
    acc[3 * (size_t)i + 2] = a.z;
}


}  // namespace gb
