// TEST INFRASTRUCTURE: compiles the product's backward arithmetic (csrc/egnn_backward_math.cuh, the same file the
// CUDA kernels include) with g++ and runs it edge by edge / node by node on the CPU, so that every formula can be
// checked against torch autograd of the oracle in this GPU-less container (tests/test_backward_math.py).
// The accumulation of the weight gradients below is the specification the kernels' tile reductions implement.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../se3-equi-graph-registration_b200/csrc/egnn_backward_math.cuh"

using namespace egspr::bwd;

namespace {
struct HostSink {
    float c[C_COUNT][32];
    template <int ID> void col(const float (&x)[32]) { std::memcpy(c[ID], x, sizeof(x)); }
};
}  // namespace

// One E_GCL layer, forward (to obtain P, Q, agg) + backward.  row/col: global node ids per edge (any order).
// edge_attr: per-edge scalar or NULL (= ea_const).  Outputs: dh_in [G][32], dx_in [G][3], gpack [7104] (+=).
// pack must be 16-byte aligned.
extern "C" void egcl_backward_host(const float *pack, int G, int E, const int32_t *row, const int32_t *col,
                                   const float *edge_attr, float ea_const, const float *h, const float *x,
                                   const float *dh_out, const float *dx_out, float *dh_in, float *dx_in,
                                   float *gpack) {
    std::vector<float> P((size_t)G * 32), Q((size_t)G * 32), agg((size_t)G * 32, 0.f);
    for (int n = 0; n < G; ++n)
        for (int o = 0; o < 32; ++o) {
            float p = 0.f, q = pack[B_BQ + o];
            for (int i = 0; i < 32; ++i) {
                p += pack[B_WPT + 32 * i + o] * h[(size_t)n * 32 + i];
                q += pack[B_WQT + 32 * i + o] * h[(size_t)n * 32 + i];
            }
            P[(size_t)n * 32 + o] = p;
            Q[(size_t)n * 32 + o] = q;
        }
    const float zero32[32] = {0}, zero3[3] = {0, 0, 0};
    HostSink sink;
    float rM[32], rDC1[32], rA1[32], rDU[32], rDPRE[32], geo[13], dxr[3], dxc[3];
    for (int e = 0; e < E; ++e) {   // forward: messages -> agg
        const int r = row[e], c = col[e];
        for (int o = 0; o < 32; ++o) rDPRE[o] = P[(size_t)r * 32 + o] + Q[(size_t)c * 32 + o];
        edge_backward<1>(pack, pack + B_WEA, x + 3 * (size_t)r, x + 3 * (size_t)c, edge_attr ? edge_attr[e] : ea_const,
                      zero32, zero3, rM, rDC1, rA1, rDU, rDPRE, geo, sink, dxr, dxc);
        for (int o = 0; o < 32; ++o) agg[(size_t)r * 32 + o] += rM[o];
    }
    std::vector<float> dagg((size_t)G * 32), dP((size_t)G * 32, 0.f), dQ((size_t)G * 32, 0.f);
    for (int n = 0; n < G; ++n) {   // node MLP backward
        float a[32], dz1[32];
        const float *hn = h + (size_t)n * 32, *an = &agg[(size_t)n * 32], *dn = dh_out + (size_t)n * 32;
        node_backward(pack, hn, an, dn, dn, a, dz1, dh_in + (size_t)n * 32, &dagg[(size_t)n * 32]);
        for (int i = 0; i < 32; ++i)
            for (int o = 0; o < 32; ++o) {
                gpack[B_WN2T + 32 * i + o] += a[i] * dn[o];
                gpack[B_WN1T + 32 * i + o] += hn[i] * dz1[o];
                gpack[B_WN1T + 32 * (32 + i) + o] += an[i] * dz1[o];
            }
        for (int o = 0; o < 32; ++o) { gpack[B_BN2 + o] += dn[o]; gpack[B_BN1 + o] += dz1[o]; }
        for (int i = 0; i < 3; ++i) dx_in[(size_t)n * 3 + i] = dx_out[(size_t)n * 3 + i];   // x_out = x + sum trans
    }
    for (int e = 0; e < E; ++e) {   // edge backward
        const int r = row[e], c = col[e];
        for (int o = 0; o < 32; ++o) rDPRE[o] = P[(size_t)r * 32 + o] + Q[(size_t)c * 32 + o];
        edge_backward<1>(pack, pack + B_WEA, x + 3 * (size_t)r, x + 3 * (size_t)c, edge_attr ? edge_attr[e] : ea_const,
                      &dagg[(size_t)r * 32], dx_out + 3 * (size_t)r, rM, rDC1, rA1, rDU, rDPRE, geo, sink, dxr, dxc);
        for (int o = 0; o < 32; ++o) { dP[(size_t)r * 32 + o] += rDPRE[o]; dQ[(size_t)c * 32 + o] += rDPRE[o]; }
        for (int i = 0; i < 3; ++i) { dx_in[(size_t)r * 3 + i] += dxr[i]; dx_in[(size_t)c * 3 + i] += dxc[i]; }
        for (int o = 0; o < 32; ++o)
            for (int i = 0; i < 32; ++i) gpack[B_WC1 + 32 * o + i] += rDC1[o] * rM[i];
        for (int hd = 0; hd < 4; ++hd)
            for (int i = 0; i < 8; ++i)
                for (int o = 0; o < 8; ++o) gpack[B_W2P + 64 * hd + 8 * i + o] += rA1[8 * hd + i] * rDU[8 * hd + o];
        for (int k = 0; k < 12; ++k)
            for (int o = 0; o < 32; ++o) gpack[B_WG + 32 * k + o] += geo[k] * rDPRE[o];
        for (int o = 0; o < 32; ++o) {
            gpack[B_WEA + o] += geo[12] * rDPRE[o];
            gpack[B_WC2 + o] += sink.c[C_DWC2][o];
            gpack[B_BC1 + o] += sink.c[C_DBC1][o];
            gpack[B_LNB + o] += sink.c[C_DLNB][o];
            gpack[B_LNG + o] += sink.c[C_DLNG][o];
            gpack[B_B2 + o] += sink.c[C_DB2][o];
        }
    }
    for (int n = 0; n < G; ++n) {   // P/Q halves of the first edge Linear back to h
        const float *hn = h + (size_t)n * 32, *p = &dP[(size_t)n * 32], *q = &dQ[(size_t)n * 32];
        linear32_backward_input(pack + B_WPT, p, dh_in + (size_t)n * 32, true);
        linear32_backward_input(pack + B_WQT, q, dh_in + (size_t)n * 32, true);
        for (int i = 0; i < 32; ++i)
            for (int o = 0; o < 32; ++o) {
                gpack[B_WPT + 32 * i + o] += hn[i] * p[o];
                gpack[B_WQT + 32 * i + o] += hn[i] * q[o];
            }
        for (int o = 0; o < 32; ++o) gpack[B_BQ + o] += q[o];
    }
}
