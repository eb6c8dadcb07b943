// Per-edge / per-node arithmetic of the E_GCL BACKWARD pass (the gradient of
// src/3dmatch_train_egnn_with_batch.py:128-289 that `loss.backward()` at 3dm:1125 produces through autograd).
//
// Host/device code without intrinsics: the CUDA kernels (egnn_backward.cu) call these functions one edge / one
// node per thread with the weight pack in shared memory, and tests/bwd_host_harness.cpp compiles the SAME file
// with g++ to check every formula against torch autograd of the oracle on the CPU (no GPU needed).
//
// The forward state of an edge is recomputed here from the layer input (h -> P,Q; x), nothing per-edge is kept
// from the forward pass.  Vectors a weight gradient needs are handed to a `Sink`:
//     sink.vec<ID>(v[32])     rows of the outer products   dW += sum_e out_e (x) in_e
//     sink.col<ID>(v[32])     rows whose column sums are bias / LayerNorm / wc2 gradients
//     sink.geo(g[13])         the 13 geometric inputs [radial, dist, dot, so3(9), edge_attr]
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EGSPR_HD __host__ __device__ __forceinline__
#else
#define EGSPR_HD inline
#endif

namespace egspr {
namespace bwd {

// layer-pack offsets (floats), identical to egspr_common.cuh / packing.py
constexpr int B_WG = 0, B_W2P = 384, B_B2 = 640, B_LNG = 672, B_LNB = 704, B_WC1 = 736, B_BC1 = 1760, B_WC2 = 1792,
              B_WN1T = 1824, B_BN1 = 3872, B_WN2T = 3904, B_BN2 = 4928, B_WPT = 4960, B_WQT = 5984, B_BQ = 7008,
              B_WEA = 7040;

enum VecId { V_M = 0, V_DC1 = 1, V_A1 = 2, V_DU = 3, V_DPRE = 4, V_COUNT = 5 };
enum ColId { C_DWC2 = 0, C_DBC1 = 1, C_DLNB = 2, C_DLNG = 3, C_DB2 = 4, C_COUNT = 5 };

EGSPR_HD float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// cross product
EGSPR_HD void cross3(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// geometry of one edge (3dm:128-181, 271-278): geo[0..11] = radial, dist, dot, so3 row-major with columns (a,b,c)
struct EdgeGeo {
    float d[3], cr[3], a[3], b[3];
    float dist, nb;     // |d|, |xr x xc|
    bool bad;           // identity frame (3dm:152-163): no gradient through the frame
};

EGSPR_HD void edge_geometry(const float *xr, const float *xc, EdgeGeo &g, float *geo) {
    for (int i = 0; i < 3; ++i) g.d[i] = xr[i] - xc[i];                        // :273
    const float radial = g.d[0] * g.d[0] + g.d[1] * g.d[1] + g.d[2] * g.d[2];  // :274
    g.dist = sqrtf(radial);                                                    // :179
    const float ia = 1.0f / (g.dist + 1e-8f);                                  // :140
    for (int i = 0; i < 3; ++i) g.a[i] = g.d[i] * ia;
    cross3(xr, xc, g.cr);                                                      // :143
    g.nb = sqrtf(g.cr[0] * g.cr[0] + g.cr[1] * g.cr[1] + g.cr[2] * g.cr[2]);
    const float ib = 1.0f / (g.nb + 1e-8f);                                    // :144
    for (int i = 0; i < 3; ++i) g.b[i] = g.cr[i] * ib;
    float c[3];
    cross3(g.a, g.b, c);                                                       // :149
    const float na2 = g.a[0] * g.a[0] + g.a[1] * g.a[1] + g.a[2] * g.a[2];
    const float nb2 = g.b[0] * g.b[0] + g.b[1] * g.b[1] + g.b[2] * g.b[2];
    const float nc2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    g.bad = na2 < 1e-12f || nb2 < 1e-12f || nc2 < 1e-12f;                      // :152-156
    geo[0] = radial; geo[1] = g.dist; geo[2] = xr[0] * xc[0] + xr[1] * xc[1] + xr[2] * xc[2];   // :180
    if (g.bad) {
        geo[3] = 1.f; geo[4] = 0.f; geo[5] = 0.f; geo[6] = 0.f; geo[7] = 1.f; geo[8] = 0.f; geo[9] = 0.f; geo[10] = 0.f; geo[11] = 1.f;
    } else {
        geo[3] = g.a[0]; geo[4] = g.b[0]; geo[5] = c[0];
        geo[6] = g.a[1]; geo[7] = g.b[1]; geo[8] = c[1];
        geo[9] = g.a[2]; geo[10] = g.b[2]; geo[11] = c[2];
    }
}

// gradient of the 12 geometric inputs w.r.t. the two endpoints.  gg[12] = d loss / d geo; gd_extra = gradient
// arriving at coord_diff from the coordinate update (trans = coord_diff * s, 3dm:264).  Norms at exactly 0 follow
// torch's subgradient convention (0).
EGSPR_HD void edge_geometry_backward(const float *xr, const float *xc, const EdgeGeo &g, const float *gg,
                                     const float *gd_extra, float *dxr, float *dxc) {
    float gd[3] = {gd_extra[0], gd_extra[1], gd_extra[2]};
    float gcr[3] = {0.f, 0.f, 0.f};
    if (!g.bad) {
        float ga[3] = {gg[3], gg[6], gg[9]}, gb[3] = {gg[4], gg[7], gg[10]};
        const float gc[3] = {gg[5], gg[8], gg[11]};
        float t[3];
        cross3(g.b, gc, t);                                  // c = a x b:  ga += b x gc, gb += gc x a
        for (int i = 0; i < 3; ++i) ga[i] += t[i];
        cross3(gc, g.a, t);
        for (int i = 0; i < 3; ++i) gb[i] += t[i];
        // a = d / (|d| + eps)
        const float sa = 1.0f / (g.dist + 1e-8f);
        const float gad = ga[0] * g.d[0] + ga[1] * g.d[1] + ga[2] * g.d[2];
        const float ka = g.dist > 0.f ? gad * sa * sa / g.dist : 0.f;
        for (int i = 0; i < 3; ++i) gd[i] += sa * ga[i] - ka * g.d[i];
        // b = cr / (|cr| + eps)
        const float sb = 1.0f / (g.nb + 1e-8f);
        const float gbc = gb[0] * g.cr[0] + gb[1] * g.cr[1] + gb[2] * g.cr[2];
        const float kb = g.nb > 0.f ? gbc * sb * sb / g.nb : 0.f;
        for (int i = 0; i < 3; ++i) gcr[i] = sb * gb[i] - kb * g.cr[i];
    }
    const float kd = (g.dist > 0.f ? gg[1] / g.dist : 0.f) + 2.0f * gg[0];   // dist = |d|, radial = |d|^2
    for (int i = 0; i < 3; ++i) gd[i] += kd * g.d[i];
    float t[3];
    cross3(xc, gcr, t);                                      // cr = xr x xc: dxr = xc x gcr, dxc = gcr x xr
    for (int i = 0; i < 3; ++i) dxr[i] = gd[i] + gg[2] * xc[i] + t[i];
    cross3(gcr, xr, t);
    for (int i = 0; i < 3; ++i) dxc[i] = -gd[i] + gg[2] * xr[i] + t[i];
}

// One edge: recompute the forward (edge_model 3dm:231-250, coord_model 3dm:262-268) and push the gradient back.
//   w        layer pack (shared memory on the device)
//   xr, xc   coordinates of row / col endpoint;  Pr = P[row], Qc = Q[col] (32 floats each);  ea = edge_attr value
//   dagg     d loss / d agg[row]  (32)           dxo = d loss / d coord_out[row]  (3)
// Outputs: dpre[32] (= gradient of P[row] and of Q[col]), dxr[3], dxc[3]; weight-gradient rows go to `sink`.
template <class Sink>
EGSPR_HD void edge_backward(const float *w, const float *xr, const float *xc, const float *Pr, const float *Qc,
                            float ea, const float *dagg, const float *dxo, Sink &sink, float *dpre, float *dxr,
                            float *dxc) {
    EdgeGeo g;
    float geo[13];
    edge_geometry(xr, xc, g, geo);
    geo[12] = ea;
    sink.geo(geo);
    // first edge Linear, P/Q factorised (bias folded in Q), + SiLU
    float a1[32], ds1[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float pre = Pr[o] + Qc[o];
#pragma unroll
        for (int k = 0; k < 12; ++k) pre = fmaf(w[B_WG + 32 * k + o], geo[k], pre);
        pre = fmaf(w[B_WEA + o], ea, pre);
        const float sg = sigmoidf_(pre);
        a1[o] = pre * sg;
        ds1[o] = sg * (1.0f + pre * (1.0f - sg));           // d silu / d pre
    }
    sink.template vec<V_A1>(a1);
    // per-head second Linear (block diagonal) + LayerNorm(32), eps 1e-5, biased variance  (:245-249)
    float uh[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) uh[o] = w[B_B2 + o];
#pragma unroll
    for (int hd = 0; hd < 4; ++hd)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int o = 0; o < 8; ++o) uh[8 * hd + o] = fmaf(w[B_W2P + 64 * hd + 8 * i + o], a1[8 * hd + i], uh[8 * hd + o]);
    float mean = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) mean += uh[o];
    mean *= (1.0f / 32.0f);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) { const float t = uh[o] - mean; var = fmaf(t, t, var); }
    const float rstd = 1.0f / sqrtf(var * (1.0f / 32.0f) + 1e-5f);
    float m[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) { uh[o] = (uh[o] - mean) * rstd; m[o] = fmaf(uh[o], w[B_LNG + o], w[B_LNB + o]); }
    sink.template vec<V_M>(m);
    // coord MLP: s = wc2 . SiLU(Wc1 m + bc1)  (:219-229);  trans = coord_diff * s  (:264)
    const float dsc = g.d[0] * dxo[0] + g.d[1] * dxo[1] + g.d[2] * dxo[2];   // d loss / d s
    float s = 0.f;
    float dc1[32];
    {
        float a2ds[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            float c1 = w[B_BC1 + o];
#pragma unroll
            for (int i = 0; i < 32; ++i) c1 = fmaf(w[B_WC1 + 32 * o + i], m[i], c1);
            const float sg = sigmoidf_(c1);
            const float a2 = c1 * sg;
            s = fmaf(w[B_WC2 + o], a2, s);
            a2ds[o] = a2 * dsc;
            dc1[o] = w[B_WC2 + o] * dsc * (sg * (1.0f + c1 * (1.0f - sg)));
        }
        sink.template col<C_DWC2>(a2ds);
    }
    sink.template vec<V_DC1>(dc1);
    sink.template col<C_DBC1>(dc1);
    // message gradient: from the node aggregate and from the coord MLP
    float dm[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dm[i] = dagg[i];
#pragma unroll
    for (int o = 0; o < 32; ++o)
#pragma unroll
        for (int i = 0; i < 32; ++i) dm[i] = fmaf(w[B_WC1 + 32 * o + i], dc1[o], dm[i]);
    sink.template col<C_DLNB>(dm);
    // LayerNorm backward
    float du[32];
    {
        float dg[32];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            dg[i] = dm[i] * uh[i];
            const float duh = dm[i] * w[B_LNG + i];
            du[i] = duh;
            s1 += duh;
            s2 = fmaf(duh, uh[i], s2);
        }
        sink.template col<C_DLNG>(dg);
        s1 *= (1.0f / 32.0f);
        s2 *= (1.0f / 32.0f);
#pragma unroll
        for (int i = 0; i < 32; ++i) du[i] = rstd * (du[i] - s1 - uh[i] * s2);
    }
    sink.template vec<V_DU>(du);
    sink.template col<C_DB2>(du);
    // second Linear backward + SiLU backward
#pragma unroll
    for (int hd = 0; hd < 4; ++hd)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = 0.f;
#pragma unroll
            for (int o = 0; o < 8; ++o) t = fmaf(w[B_W2P + 64 * hd + 8 * i + o], du[8 * hd + o], t);
            dpre[8 * hd + i] = t * ds1[8 * hd + i];
        }
    {
        float dp[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) dp[o] = dpre[o];
        sink.template vec<V_DPRE>(dp);
    }
    // geometry backward (+ the coordinate update's own use of coord_diff)
    float gg[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float t = 0.f;
#pragma unroll
        for (int o = 0; o < 32; ++o) t = fmaf(w[B_WG + 32 * k + o], dpre[o], t);
        gg[k] = t;
    }
    const float gde[3] = {s * dxo[0], s * dxo[1], s * dxo[2]};
    edge_geometry_backward(xr, xc, g, gg, gde, dxr, dxc);
}

// node_model backward (3dm:252-260): out = h + Wn2 SiLU(Wn1 [h|agg] + bn1) + bn2.
//   dout = d loss / d h_out[n].  Outputs: dh (the part that does not go through P/Q), dagg, and the rows the
//   weight gradients need: a (SiLU output), dz1.
EGSPR_HD void node_backward(const float *w, const float *h, const float *agg, const float *dout, float *dh, float *dagg,
                            float *a, float *dz1) {
    float ds[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float z = w[B_BN1 + o];
#pragma unroll
        for (int i = 0; i < 32; ++i) z = fmaf(w[B_WN1T + 32 * i + o], h[i], z);
#pragma unroll
        for (int i = 0; i < 32; ++i) z = fmaf(w[B_WN1T + 32 * (32 + i) + o], agg[i], z);
        const float sg = sigmoidf_(z);
        a[o] = z * sg;
        ds[o] = sg * (1.0f + z * (1.0f - sg));
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float t = 0.f;
#pragma unroll
        for (int o = 0; o < 32; ++o) t = fmaf(w[B_WN2T + 32 * i + o], dout[o], t);
        dz1[i] = t * ds[i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float t = dout[i], u = 0.f;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            t = fmaf(w[B_WN1T + 32 * i + o], dz1[o], t);
            u = fmaf(w[B_WN1T + 32 * (32 + i) + o], dz1[o], u);
        }
        dh[i] = t;
        dagg[i] = u;
    }
}

// y = W x (+ b) with W stored transposed ([in][out]):  dx[i] (+)= sum_o wt[32 i + o] dy[o]
EGSPR_HD void linear32_backward_input(const float *wt, const float *dy, float *dx, bool accumulate) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float t = accumulate ? dx[i] : 0.f;
#pragma unroll
        for (int o = 0; o < 32; ++o) t = fmaf(wt[32 * i + o], dy[o], t);
        dx[i] = t;
    }
}

}  // namespace bwd
}  // namespace egspr
