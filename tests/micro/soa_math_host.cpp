// Host check (developer tool + CPU test helper) of the pair-packed correlation arithmetic in
// torchpiv_b200/csrc/piv_soa_math.cuh: emulates the H lanes of one window with plain arrays for the
// shared-memory exchanges and compares the result with the direct circular cross-correlation.
//   g++ -std=c++17 -O1 -I/usr/local/cuda/include -o soa_math_host soa_math_host.cpp && ./soa_math_host
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../torchpiv_b200/csrc/piv_soa_math.cuh"

using namespace pivb200;

template <int W>
static double run_case(unsigned seed, bool drop_dc) {
    constexpr int H = W / 2;
    using M = SoaMath<W>;
    std::vector<float> a(W * W), b(W * W);
    srand(seed);
    for (int i = 0; i < W * W; ++i) { a[i] = float(rand() % 256); b[i] = float(rand() % 256); }
    // quads [row pair][column]: (re pair, im pair)
    std::vector<float2> X(2 * H * H), Q(2 * H * H);
    std::vector<std::vector<float2>> park(H, std::vector<float2>(W));
    for (int frame = 0; frame < 2; ++frame) {
        const std::vector<float>& f = frame ? b : a;
        for (int l = 0; l < H; ++l) {
            float2 x[W];
            for (int j = 0; j < W; ++j) x[j] = make_float2(f[(2 * l) * W + j], f[(2 * l + 1) * W + j]);
            M::row_forward(x);
            for (int c = 0; c < H; ++c) {
                X[2 * (l * H + c)] = x[2 * M::pos(c)];
                X[2 * (l * H + c) + 1] = x[2 * M::pos(c) + 1];
            }
        }
        std::vector<std::vector<float2>> col(H, std::vector<float2>(W));
        for (int c = 0; c < H; ++c) {
            float2 x[W];
            for (int t = 0; t < H; ++t) { x[2 * t] = X[2 * (t * H + c)]; x[2 * t + 1] = X[2 * (t * H + c) + 1]; }
            M::col_forward(x);
            for (int i = 0; i < W; ++i) col[c][i] = x[i];
        }
        if (frame == 0) { park = col; continue; }
        // product
        for (int c = 0; c < H; ++c) {
            float2 xa[W], xb[W];
            for (int i = 0; i < W; ++i) { xa[i] = park[c][i]; xb[i] = col[c][i]; }
            if (c > 0) {
                M::product(xa, xb);
            } else {
                // column 0 = two real columns (bins 0 and W/2) packed as one complex column
                float2 sA[W], sB[W], sV[W];
                for (int q = 0; q < H; ++q) {
                    const int e = M::pos(q);
                    sA[q] = make_float2(xa[2 * e].x, xa[2 * e + 1].x);
                    sA[q + H] = make_float2(xa[2 * e].y, xa[2 * e + 1].y);
                    sB[q] = make_float2(xb[2 * e].x, xb[2 * e + 1].x);
                    sB[q + H] = make_float2(xb[2 * e].y, xb[2 * e + 1].y);
                }
                for (int l = 0; l < H; ++l) {
                    if (l == 0) {
                        const float p00 = drop_dc ? 0.f : sA[0].x * sB[0].x;
                        sV[0] = make_float2(p00, sA[0].y * sB[0].y);
                        sV[H] = make_float2(sA[H].x * sB[H].x, sA[H].y * sB[H].y);
                    } else {
                        const float2 Aq = sA[l], An = sA[W - l], Bq = sB[l], Bn = sB[W - l];
                        const float2 a0 = make_float2(Aq.x + An.x, Aq.y - An.y), ah = make_float2(Aq.y + An.y, An.x - Aq.x);
                        const float2 b0 = make_float2(0.25f * (Bq.x + Bn.x), 0.25f * (Bq.y - Bn.y));
                        const float2 bh = make_float2(0.25f * (Bq.y + Bn.y), 0.25f * (Bn.x - Bq.x));
                        const float2 P0 = make_float2(a0.x * b0.x + a0.y * b0.y, a0.x * b0.y - a0.y * b0.x);
                        const float2 Ph = make_float2(ah.x * bh.x + ah.y * bh.y, ah.x * bh.y - ah.y * bh.x);
                        sV[l] = make_float2(P0.x - Ph.y, P0.y + Ph.x);
                        sV[W - l] = make_float2(P0.x + Ph.y, Ph.x - P0.y);
                    }
                }
                for (int q = 0; q < H; ++q) {
                    const int e = M::pos(q);
                    xb[2 * e] = make_float2(sV[q].x, sV[q + H].x);
                    xb[2 * e + 1] = make_float2(sV[q].y, sV[q + H].y);
                }
            }
            M::col_inverse(xb);                    // result swapped: real pair in the odd slot
            for (int m = 0; m < H; ++m) {
                Q[2 * (m * H + c)] = xb[2 * M::pos(m) + 1];
                Q[2 * (m * H + c) + 1] = xb[2 * M::pos(m)];
            }
        }
    }
    std::vector<double> got(W * W);
    for (int l = 0; l < H; ++l) {
        float2 x[W];
        for (int c = 0; c < H; ++c) { x[2 * c + 1] = Q[2 * (l * H + c)]; x[2 * c] = Q[2 * (l * H + c) + 1]; }   // swapped
        M::row_inverse(x);
        for (int m = 0; m < H; ++m) {
            const float2 ev = x[2 * M::pos(m) + 1], od = x[2 * M::pos(m)];
            got[(2 * l) * W + 2 * m] = ev.x; got[(2 * l + 1) * W + 2 * m] = ev.y;
            got[(2 * l) * W + 2 * m + 1] = od.x; got[(2 * l + 1) * W + 2 * m + 1] = od.y;
        }
    }
    double sa = 0, sb = 0;
    for (int i = 0; i < W * W; ++i) { sa += a[i]; sb += b[i]; }
    double worst = 0, scale = 0;
    for (int sy = 0; sy < W; ++sy)
        for (int sx = 0; sx < W; ++sx) {
            double acc = 0;
            for (int y = 0; y < W; ++y)
                for (int x = 0; x < W; ++x) acc += double(a[y * W + x]) * b[((y + sy) % W) * W + (x + sx) % W];
            if (drop_dc) acc -= sa * sb / (W * W);
            const double g = got[sy * W + sx] / (4.0 * W * W);
            worst = std::fmax(worst, std::fabs(g - acc));
            scale = std::fmax(scale, std::fabs(acc));
        }
    return worst / scale;
}

int main() {
    int bad = 0;
    for (int s = 1; s <= 3; ++s) {
        const double e16 = run_case<16>(s, s & 1), e32 = run_case<32>(s, s & 1), e64 = run_case<64>(s, s & 1);
        printf("seed %d: rel err W=16 %.3g  W=32 %.3g  W=64 %.3g\n", s, e16, e32, e64);
        bad += (e16 > 2e-6) + (e32 > 2e-6) + (e64 > 2e-6);
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}
