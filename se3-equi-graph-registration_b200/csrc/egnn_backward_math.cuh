// Per-edge / per-node arithmetic of the E_GCL BACKWARD pass (the gradient of
// src/3dmatch_train_egnn_with_batch.py:128-289 that `loss.backward()` at 3dm:1125 produces through autograd).
//
// Host/device code without intrinsics: the CUDA kernels (egnn_backward.cu) call these functions one edge / one
// node per thread with the weight pack in shared memory, and tests/bwd_host_harness.cpp compiles the SAME file
// with g++ to check every formula against torch autograd of the oracle on the CPU (no GPU needed).
//
// The forward state of an edge is recomputed here from the layer input (h -> P,Q; x), nothing per-edge is kept
// from the forward pass.  The 32-vectors the weight gradients need (rows of the outer products
// dW += sum_e out_e (x) in_e) are written to caller-provided ROWS -- on the device the thread's private rows of the
// shared-memory tile the CTA reduces afterwards, on the host plain arrays.  The rows double as the working storage
// of the loops: every large loop is rolled over an index that addresses a row (or the weights) and unrolled over
// an index that addresses registers, so the code stays small and no register array is indexed dynamically.
// Vectors whose column sums are bias / LayerNorm / wc2 gradients go to `sink.col<ID>(v[32])`.
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

struct alignas(16) F4 { float x, y, z, w; };
EGSPR_HD F4 ld4(const float *p) { return *reinterpret_cast<const F4 *>(p); }     // p 16-byte aligned (weights)

// (c0, c1) += (a0, a1) * (b0, b1), each lane rounded like fmaf: one FFMA2 issue slot on sm_100
EGSPR_HD void fma2(float &c0, float &c1, float a0, float a1, float b0, float b1) {
#ifdef __CUDA_ARCH__
    asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rd, {%0, %1};\n"
        "fma.rn.f32x2 rd, ra, rb, rd;\nmov.b64 {%0, %1}, rd;\n}"
        : "+f"(c0), "+f"(c1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
#else
    c0 = fmaf(a0, b0, c0);
    c1 = fmaf(a1, b1, c1);
#endif
}

EGSPR_HD float sigmoidf_(float v) {
#ifdef __CUDA_ARCH__
    return __fdividef(1.0f, 1.0f + __expf(-v));      // MUFU.EX2 + MUFU.RCP, ~2 ulp: far inside the gradient tolerance
#else
    return 1.0f / (1.0f + expf(-v));
#endif
}

// sqrt / reciprocal: MUFU approximations on the device (~2 ulp, far inside the gradient tolerance), IEEE on the host
EGSPR_HD float sqrt_(float v) {
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return sqrtf(v);
#endif
}
EGSPR_HD float rcp_(float v) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return 1.0f / v;
#endif
}

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
    g.dist = sqrt_(radial);                                                    // :179
    const float ia = rcp_(g.dist + 1e-8f);                                     // :140
    for (int i = 0; i < 3; ++i) g.a[i] = g.d[i] * ia;
    cross3(xr, xc, g.cr);                                                      // :143
    g.nb = sqrt_(g.cr[0] * g.cr[0] + g.cr[1] * g.cr[1] + g.cr[2] * g.cr[2]);
    const float ib = rcp_(g.nb + 1e-8f);                                       // :144
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
        const float sa = rcp_(g.dist + 1e-8f);
        const float gad = ga[0] * g.d[0] + ga[1] * g.d[1] + ga[2] * g.d[2];
        const float ka = g.dist > 0.f ? gad * sa * sa * rcp_(g.dist) : 0.f;
        for (int i = 0; i < 3; ++i) gd[i] += sa * ga[i] - ka * g.d[i];
        // b = cr / (|cr| + eps)
        const float sb = rcp_(g.nb + 1e-8f);
        const float gbc = gb[0] * g.cr[0] + gb[1] * g.cr[1] + gb[2] * g.cr[2];
        const float kb = g.nb > 0.f ? gbc * sb * sb * rcp_(g.nb) : 0.f;
        for (int i = 0; i < 3; ++i) gcr[i] = sb * gb[i] - kb * g.cr[i];
    }
    const float kd = (g.dist > 0.f ? gg[1] * rcp_(g.dist) : 0.f) + 2.0f * gg[0];   // dist = |d|, radial = |d|^2
    for (int i = 0; i < 3; ++i) gd[i] += kd * g.d[i];
    float t[3];
    cross3(xc, gcr, t);                                      // cr = xr x xc: dxr = xc x gcr, dxc = gcr x xr
    for (int i = 0; i < 3; ++i) dxr[i] = gd[i] + gg[2] * xc[i] + t[i];
    cross3(gcr, xr, t);
    for (int i = 0; i < 3; ++i) dxc[i] = -gd[i] + gg[2] * xr[i] + t[i];
}

// One edge: recompute the forward (edge_model 3dm:231-250, coord_model 3dm:262-268) and push the gradient back.
//   w        layer pack, 16-byte aligned (shared memory on the device; only the edge part [0, 1824) is read);
//            wea = the edge_attr column (pack + B_WEA)
//   xr, xc   coordinates of row / col endpoint;  ea = edge_attr value
//   dagg     d loss / d agg[row]  (32 floats)    dxo = d loss / d coord_out[row]  (3)
//   rows     rM, rA1: contiguous 32-float rows; rDC1, rDU, rDPRE: element j lives at [j * OS] (the device keeps these
//            three feature-major, OS = tile size, so that the tile reduction reads 4 edges per 128-bit load).
//            rDPRE holds P[row] + Q[col] on entry.  On return: rM = message m, rDC1 = d loss / d (coord_mlp.0 output),
//            rA1 = SiLU output of the first edge Linear, rDU = d loss / d (LayerNorm input), rDPRE = d loss / d (first
//            Linear output) = the gradient of P[row] and of Q[col];  geo[13] = [radial, dist, dot, so3(9), edge_attr]
// Outputs: dxr[3], dxc[3] (coordinate gradients of the two endpoints).
template <int OS, class Sink>
EGSPR_HD void edge_backward(const float *w, const float *wea, const float *xr, const float *xc, float ea,
                            const float *dagg, const float *dxo, float *rM, float *rDC1, float *rA1, float *rDU,
                            float *rDPRE, float *geo, Sink &sink, float *dxr, float *dxc) {
    EdgeGeo g;
    edge_geometry(xr, xc, g, geo);
    geo[12] = ea;
    float gk[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) gk[k] = geo[k];
    // first edge Linear, P/Q factorised (bias folded in Q), + SiLU.  rDPRE: pq -> d silu / d pre
#pragma unroll 2
    for (int o4 = 0; o4 < 32; o4 += 4) {
        const F4 we = ld4(wea + o4);
        float p0 = fmaf(we.x, ea, rDPRE[(o4) * OS]), p1 = fmaf(we.y, ea, rDPRE[(o4 + 1) * OS]), p2 = fmaf(we.z, ea, rDPRE[(o4 + 2) * OS]),
              p3 = fmaf(we.w, ea, rDPRE[(o4 + 3) * OS]);
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const F4 wv = ld4(w + B_WG + 32 * k + o4);
            fma2(p0, p1, wv.x, wv.y, gk[k], gk[k]); fma2(p2, p3, wv.z, wv.w, gk[k], gk[k]);
        }
        const float pr[4] = {p0, p1, p2, p3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float sg = sigmoidf_(pr[q]);
            rA1[o4 + q] = pr[q] * sg;
            rDPRE[(o4 + q) * OS] = sg * (1.0f + pr[q] * (1.0f - sg));
        }
    }
    // per-head second Linear (block diagonal) + LayerNorm(32), eps 1e-5, biased variance  (:245-249)
    float uh[32];
#pragma unroll
    for (int o4 = 0; o4 < 32; o4 += 4) {
        const F4 b = ld4(w + B_B2 + o4);
        uh[o4] = b.x; uh[o4 + 1] = b.y; uh[o4 + 2] = b.z; uh[o4 + 3] = b.w;
    }
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) {
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const float av = rA1[8 * hd + i];
            const F4 w0 = ld4(w + B_W2P + 64 * hd + 8 * i), w1 = ld4(w + B_W2P + 64 * hd + 8 * i + 4);
            float *u = uh + 8 * hd;
            fma2(u[0], u[1], w0.x, w0.y, av, av); fma2(u[2], u[3], w0.z, w0.w, av, av);
            fma2(u[4], u[5], w1.x, w1.y, av, av); fma2(u[6], u[7], w1.z, w1.w, av, av);
        }
    }
    float mean = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) mean += uh[o];
    mean *= (1.0f / 32.0f);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) { const float t = uh[o] - mean; var = fmaf(t, t, var); }
    const float rstd = 1.0f / sqrtf(var * (1.0f / 32.0f) + 1e-5f);
    // coord MLP: s = wc2 . SiLU(Wc1 m + bc1)  (:219-229);  trans = coord_diff * s  (:264)
    const float dsc = g.d[0] * dxo[0] + g.d[1] * dxo[1] + g.d[2] * dxo[2];   // d loss / d s
    float s = 0.f;
    {
        float m[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            uh[o] = (uh[o] - mean) * rstd;
            m[o] = fmaf(uh[o], w[B_LNG + o], w[B_LNB + o]);
            rM[o] = m[o];
        }
#pragma unroll 2
        for (int o = 0; o < 32; ++o) {
            float c0 = w[B_BC1 + o], c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;   // 4 independent FFMA2 chains
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                const F4 w0 = ld4(w + B_WC1 + 32 * o + i), w1 = ld4(w + B_WC1 + 32 * o + i + 4);
                fma2(c0, c1, w0.x, w0.y, m[i], m[i + 1]); fma2(c2, c3, w0.z, w0.w, m[i + 2], m[i + 3]);
                fma2(c4, c5, w1.x, w1.y, m[i + 4], m[i + 5]); fma2(c6, c7, w1.z, w1.w, m[i + 6], m[i + 7]);
            }
            c0 += c2; c1 += c3; c4 += c6; c5 += c7; c0 += c4; c1 += c5;
            const float c = c0 + c1;
            const float sg = sigmoidf_(c);
            const float a2 = c * sg;
            const float wc = w[B_WC2 + o];
            s = fmaf(wc, a2, s);
            rDU[(o) * OS] = a2 * dsc;                                              // scratch: rows of the wc2 gradient
            rDC1[(o) * OS] = wc * dsc * (sg * (1.0f + c * (1.0f - sg)));
        }
    }
    {
        float v[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) v[o] = rDU[(o) * OS];
        sink.template col<C_DWC2>(v);
#pragma unroll
        for (int o = 0; o < 32; ++o) v[o] = rDC1[(o) * OS];
        sink.template col<C_DBC1>(v);
    }
    // message gradient: from the node aggregate and from the coord MLP
    float dm[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dm[i] = dagg[i];
#pragma unroll 1
    for (int o = 0; o < 32; ++o) {
        const float dc = rDC1[(o) * OS];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const F4 wv = ld4(w + B_WC1 + 32 * o + i);
            fma2(dm[i], dm[i + 1], wv.x, wv.y, dc, dc); fma2(dm[i + 2], dm[i + 3], wv.z, wv.w, dc, dc);
        }
    }
    sink.template col<C_DLNB>(dm);
    // LayerNorm backward
    {
        float dg[32];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            dg[i] = dm[i] * uh[i];
            dm[i] *= w[B_LNG + i];
            s1 += dm[i];
            s2 = fmaf(dm[i], uh[i], s2);
        }
        sink.template col<C_DLNG>(dg);
        s1 *= (1.0f / 32.0f);
        s2 *= (1.0f / 32.0f);
#pragma unroll
        for (int i = 0; i < 32; ++i) { dm[i] = rstd * (dm[i] - s1 - uh[i] * s2); rDU[(i) * OS] = dm[i]; }   // dm is now du
    }
    sink.template col<C_DB2>(dm);
    // second Linear backward + SiLU backward.  rDPRE: d silu / d pre -> dpre
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) {
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const F4 w0 = ld4(w + B_W2P + 64 * hd + 8 * i), w1 = ld4(w + B_W2P + 64 * hd + 8 * i + 4);
            const float *du = dm + 8 * hd;
            float t0 = 0.f, t1 = 0.f;
            fma2(t0, t1, w0.x, w0.y, du[0], du[1]); fma2(t0, t1, w0.z, w0.w, du[2], du[3]);
            fma2(t0, t1, w1.x, w1.y, du[4], du[5]); fma2(t0, t1, w1.z, w1.w, du[6], du[7]);
            rDPRE[(8 * hd + i) * OS] *= (t0 + t1);
        }
    }
    // geometry backward (+ the coordinate update's own use of coord_diff)
    float gg[12], gh[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { gg[k] = 0.f; gh[k] = 0.f; }
#pragma unroll 1
    for (int o4 = 0; o4 < 32; o4 += 4) {
        const float d0 = rDPRE[(o4) * OS], d1 = rDPRE[(o4 + 1) * OS], d2 = rDPRE[(o4 + 2) * OS], d3 = rDPRE[(o4 + 3) * OS];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const F4 wv = ld4(w + B_WG + 32 * k + o4);
            fma2(gg[k], gh[k], wv.x, wv.y, d0, d1); fma2(gg[k], gh[k], wv.z, wv.w, d2, d3);
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) gg[k] += gh[k];
    const float gde[3] = {s * dxo[0], s * dxo[1], s * dxo[2]};
    edge_geometry_backward(xr, xc, g, gg, gde, dxr, dxc);
}

// node_model backward (3dm:252-260): out = h + Wn2 SiLU(Wn1 [h|agg] + bn1) + bn2.
//   w        layer pack (16-byte aligned); rH, rAgg: rows holding h[n], agg[n]; dout[32] = d loss / d h_out[n]
//            (registers) with a copy in the row rDout
//   rows     rA <- SiLU output, rZ <- dz1 (gradient at the first Linear's output)
//   dh[32] <- the part of d loss / d h[n] that does not go through P/Q, dagg[32] <- d loss / d agg[n]
//            (written one element at a time: pointers to global memory on the device)
EGSPR_HD void node_backward(const float *w, const float *rH, const float *rAgg, const float *dout, const float *rDout,
                            float *rA, float *rZ, float *dh, float *dagg) {
    float z[32];
#pragma unroll
    for (int o4 = 0; o4 < 32; o4 += 4) {
        const F4 b = ld4(w + B_BN1 + o4);
        z[o4] = b.x; z[o4 + 1] = b.y; z[o4 + 2] = b.z; z[o4 + 3] = b.w;
    }
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
        const float v = i < 32 ? rH[i] : rAgg[i - 32];
#pragma unroll
        for (int o4 = 0; o4 < 32; o4 += 4) {
            const F4 wv = ld4(w + B_WN1T + 32 * i + o4);
            fma2(z[o4], z[o4 + 1], wv.x, wv.y, v, v); fma2(z[o4 + 2], z[o4 + 3], wv.z, wv.w, v, v);
        }
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const float sg = sigmoidf_(z[o]);
        rA[o] = z[o] * sg;
        rZ[o] = sg * (1.0f + z[o] * (1.0f - sg));           // d silu / d z, replaced by dz1 below
    }
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int o4 = 0; o4 < 32; o4 += 4) {
            const F4 wv = ld4(w + B_WN2T + 32 * i + o4);
            fma2(t0, t1, wv.x, wv.y, dout[o4], dout[o4 + 1]); fma2(t0, t1, wv.z, wv.w, dout[o4 + 2], dout[o4 + 3]);
        }
        rZ[i] *= (t0 + t1);
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) z[o] = rZ[o];               // dz1 in registers
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int o4 = 0; o4 < 32; o4 += 4) {
            const F4 wv = ld4(w + B_WN1T + 32 * i + o4);
            fma2(t0, t1, wv.x, wv.y, z[o4], z[o4 + 1]); fma2(t0, t1, wv.z, wv.w, z[o4 + 2], z[o4 + 3]);
        }
        if (i < 32) dh[i] = rDout[i] + (t0 + t1);
        else dagg[i - 32] = t0 + t1;
    }
}

// y = W x (+ b) with W stored transposed ([in][out], 16-byte aligned):  dx[i] (+)= sum_o wt[32 i + o] dy[o];
// dy in registers, dx written one element at a time
EGSPR_HD void linear32_backward_input(const float *wt, const float *dy, float *dx, bool accumulate) {
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int o4 = 0; o4 < 32; o4 += 4) {
            const F4 wv = ld4(wt + 32 * i + o4);
            fma2(t0, t1, wv.x, wv.y, dy[o4], dy[o4 + 1]); fma2(t0, t1, wv.z, wv.w, dy[o4 + 2], dy[o4 + 3]);
        }
        dx[i] = (accumulate ? dx[i] : 0.f) + (t0 + t1);
    }
}

}  // namespace bwd
}  // namespace egspr
