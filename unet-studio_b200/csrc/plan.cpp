#include "plan.h"

#include <cstring>

namespace u3d {

int choose_kc(int c0p, int c1p) {
    auto ok = [&](int k) { return c0p % k == 0 && c1p % k == 0; };
    if (ok(64)) return 64;
    if (ok(32)) return 32;
    return 16;
}

// N tile: minimises (waves of 148 CTAs) x (time of one M=128,K=16 MMA), the MMA costing ~max(45, N/2) clk (tools/mma_bench.cu:
// the A-operand read from shared memory makes every N <= 64 equally expensive).  The deep levels have only 2..75 M tiles, so a
// full-width N tile would leave most of the 148 SMs idle; ties go to the wider tile (less re-reading of A through L2).
void choose_ntile(int np, long long m_voxels, int& ntile, int& ntiles) {
    const long long mt = (m_voxels + 119) / 120;
    int best = 16;
    long long best_cost = -1;
    for (int nt = 16; nt <= (np < 256 ? np : 256); nt += 16) {
        if (np % nt) continue;
        const long long items = mt * (np / nt);
        const long long cost = ((items + 147) / 148) * (nt / 2 > 45 ? nt / 2 : 45);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = nt; }
    }
    ntile = best;
    ntiles = np / best;
}

size_t pack_bytes(const PackDesc& d) {
    if (d.banded) return pack_bytes_band(d);
    return size_t(d.ntaps) * (d.nch[0] + d.nch[1]) * d.ntiles * d.ntile * d.kc * 2;
}

static void zero_problem(ConvProblem& P) { std::memset(&P, 0, sizeof(P)); }

static void init_pack(PackDesc& k) {
    std::memset(&k, 0, sizeof(k));
}

void plan_forward(const LayerGeom& g, std::vector<ConvProblem>& probs, std::vector<PackDesc>& packs, int& kc, int force_kc) {
    probs.clear();
    packs.clear();
    const int c0p = pad16(g.cin[0]), c1p = g.cin[1] ? pad16(g.cin[1]) : 0;
    kc = force_kc ? force_kc : choose_kc(c0p, c1p);
    int ntile, ntiles;
    choose_ntile(pad16(g.cout), (g.transposed ? 8LL : 1LL) * (g.transposed ? g.in_d : g.out_d) * (g.transposed ? g.in_h : g.out_h) *
                                    (g.transposed ? g.in_w : g.out_w), ntile, ntiles);
    if (!g.transposed) {
        ConvProblem P;
        zero_problem(P);
        PackDesc K;
        init_pack(K);
        P.c0p = c0p; P.c1p = c1p;
        P.nch0 = c0p / kc; P.nch1 = c1p / kc;
        P.in_d = g.in_d; P.in_h = g.in_h; P.in_w = g.in_w;
        P.istride = g.stride;
        P.ntaps = g.ks * g.ks * g.ks;
        const int pad = (g.ks - 1) / 2;
        int t = 0;
        for (int kz = 0; kz < g.ks; ++kz)
            for (int ky = 0; ky < g.ks; ++ky)
                for (int kx = 0; kx < g.ks; ++kx, ++t) {
                    P.taps[t] = ConvTap{int8_t(kz - pad), int8_t(ky - pad), int8_t(kx - pad), 0};
                    K.tap_ref[t] = t;
                }
        P.od = g.out_d; P.oh = g.out_h; P.ow = g.out_w;
        P.OD = g.out_d; P.OH = g.out_h; P.OW = g.out_w;
        P.ostep = 1;
        P.dst_cp = pad16(g.cout);
        P.ntile = ntile; P.ntiles = ntiles; P.n_real = g.cout;
        K.dimA = g.cout; K.dimB = g.cin[0] + g.cin[1]; K.ktaps = P.ntaps;
        K.n_is_A = 1; K.n_off = 0; K.n_real = g.cout; K.ntile = ntile; K.ntiles = ntiles;
        K.k_off[0] = 0; K.k_real[0] = g.cin[0]; K.nch[0] = P.nch0;
        K.k_off[1] = g.cin[0]; K.k_real[1] = g.cin[1]; K.nch[1] = P.nch1;
        K.kc = kc; K.ntaps = P.ntaps;
        if (force_kc == 16 && g.ks == 3 && g.stride == 1 && ntiles == 1 &&
            conv_band_wants(c0p + c1p, pad16(g.cout), 1LL * g.out_d * g.out_h * g.out_w)) {
            P.banded = 1; K.banded = 1; K.band_co = pad16(g.cout);
            std::memcpy(K.band_taps, P.taps, sizeof(K.band_taps));
        } else if (force_kc == 16 && g.ks == 3 && g.stride == 1 && ntiles == 1 && c0p == 32 && c1p == 32 &&
                   conv_band_wants(32, pad16(g.cout), 1LL * g.out_d * g.out_h * g.out_w)) {
            // concat of 32 + 32 channels: two banded launches, one per source (see ConvProblem::band_pass)
            P.banded = 1; K.banded = 1; K.band_co = pad16(g.cout);
            std::memcpy(K.band_taps, P.taps, sizeof(K.band_taps));
            ConvProblem P2 = P;
            PackDesc K2 = K;
            P.band_pass = 1; P2.band_pass = 2;
            K.nch[1] = 0; K.k_real[1] = 0;
            K2.k_off[0] = g.cin[0]; K2.k_real[0] = g.cin[1]; K2.nch[0] = P.nch1; K2.nch[1] = 0; K2.k_real[1] = 0;
            probs.push_back(P);
            packs.push_back(K);
            probs.push_back(P2);
            packs.push_back(K2);
            return;
        }
        probs.push_back(P);
        packs.push_back(K);
    } else if (conv_tma_available()) {
        // conv_transpose k2 s2 as ONE problem: a 1x1 conv with N = 8 parity blocks of Cout, scattered by the epilogue (no overlap
        // between the 8 taps, so nothing is summed across blocks)
        const int cp = pad16(g.cout);
        choose_ntile(8 * cp, 1LL * g.in_d * g.in_h * g.in_w, ntile, ntiles);
        ConvProblem P;
        zero_problem(P);
        PackDesc K;
        init_pack(K);
        P.c0p = c0p; P.c1p = 0;
        P.nch0 = c0p / kc; P.nch1 = 0;
        P.in_d = g.in_d; P.in_h = g.in_h; P.in_w = g.in_w;
        P.istride = 1;
        P.ntaps = 1;
        P.taps[0] = ConvTap{0, 0, 0, 0};
        P.od = g.in_d; P.oh = g.in_h; P.ow = g.in_w;
        P.OD = g.out_d; P.OH = g.out_h; P.OW = g.out_w;
        P.ostep = 2;
        P.dst_cp = cp;
        P.ntile = ntile; P.ntiles = ntiles; P.n_real = 8 * cp;
        P.shuffle_cp = cp; P.shuffle_nreal = g.cout;
        K.dimA = g.cin[0]; K.dimB = g.cout; K.ktaps = 8;
        K.n_is_A = 0; K.n_off = 0; K.n_real = g.cout; K.ntile = ntile; K.ntiles = ntiles;
        K.k_off[0] = 0; K.k_real[0] = g.cin[0]; K.nch[0] = P.nch0;
        K.kc = kc; K.ntaps = 1;
        K.stack_cp = cp;
        for (int par = 0; par < 8; ++par) K.stack_ref[0][par] = (signed char)par;
        probs.push_back(P);
        packs.push_back(K);
    } else {
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    ConvProblem P;
                    zero_problem(P);
                    PackDesc K;
                    init_pack(K);
                    P.c0p = c0p; P.c1p = 0;
                    P.nch0 = c0p / kc; P.nch1 = 0;
                    P.in_d = g.in_d; P.in_h = g.in_h; P.in_w = g.in_w;
                    P.istride = 1;
                    P.ntaps = 1;
                    P.taps[0] = ConvTap{0, 0, 0, 0};
                    P.od = g.in_d; P.oh = g.in_h; P.ow = g.in_w;
                    P.OD = g.out_d; P.OH = g.out_h; P.OW = g.out_w;
                    P.ostep = 2; P.ooff_z = a; P.ooff_y = b; P.ooff_x = c;
                    P.dst_cp = pad16(g.cout);
                    P.ntile = ntile; P.ntiles = ntiles; P.n_real = g.cout;
                    K.dimA = g.cin[0]; K.dimB = g.cout; K.ktaps = 8;
                    K.n_is_A = 0; K.n_off = 0; K.n_real = g.cout; K.ntile = ntile; K.ntiles = ntiles;
                    K.k_off[0] = 0; K.k_real[0] = g.cin[0]; K.nch[0] = P.nch0;
                    K.kc = kc; K.ntaps = 1; K.tap_ref[0] = (a * 2 + b) * 2 + c;
                    probs.push_back(P);
                    packs.push_back(K);
                }
    }
}

void plan_dgrad(const LayerGeom& g, int src, std::vector<ConvProblem>& probs, std::vector<PackDesc>& packs, int& kc, int force_kc) {
    probs.clear();
    packs.clear();
    const int coutp = pad16(g.cout);
    kc = force_kc ? force_kc : choose_kc(coutp, 0);
    const int n_real = g.cin[src];
    int ntile, ntiles;
    choose_ntile(pad16(n_real), 1LL * g.in_d * g.in_h * g.in_w, ntile, ntiles);   // dx lattice (all parity problems together)
    const int n_off = src ? g.cin[0] : 0;
    auto base = [&](ConvProblem& P, PackDesc& K) {
        zero_problem(P);
        init_pack(K);
        P.c0p = coutp; P.nch0 = coutp / kc;
        P.in_d = g.out_d; P.in_h = g.out_h; P.in_w = g.out_w;  // gathered tensor = dy
        P.OD = g.in_d; P.OH = g.in_h; P.OW = g.in_w;            // destination = dx
        P.dst_cp = pad16(n_real);
        P.ntile = ntile; P.ntiles = ntiles; P.n_real = n_real;
        K.ntile = ntile; K.ntiles = ntiles; K.n_real = n_real; K.n_off = n_off;
        K.k_off[0] = 0; K.k_real[0] = g.cout; K.nch[0] = P.nch0;
        K.kc = kc;
    };
    if (!g.transposed && g.stride == 1) {
        ConvProblem P;
        PackDesc K;
        base(P, K);
        const int pad = (g.ks - 1) / 2;
        P.istride = 1;
        P.ntaps = g.ks * g.ks * g.ks;
        int t = 0;
        for (int kz = 0; kz < g.ks; ++kz)
            for (int ky = 0; ky < g.ks; ++ky)
                for (int kx = 0; kx < g.ks; ++kx, ++t) {
                    P.taps[t] = ConvTap{int8_t(pad - kz), int8_t(pad - ky), int8_t(pad - kx), 0};
                    K.tap_ref[t] = t;
                }
        P.od = g.in_d; P.oh = g.in_h; P.ow = g.in_w;
        P.ostep = 1;
        K.dimA = g.cout; K.dimB = g.cin[0] + g.cin[1]; K.ktaps = P.ntaps; K.n_is_A = 0; K.ntaps = P.ntaps;
        if (force_kc == 16 && g.ks == 3 && ntiles == 1 && conv_band_wants(coutp, pad16(n_real), 1LL * g.in_d * g.in_h * g.in_w)) {
            P.banded = 1; K.banded = 1; K.band_co = pad16(n_real);
            std::memcpy(K.band_taps, P.taps, sizeof(K.band_taps));
        }
        probs.push_back(P);
        packs.push_back(K);
    } else if (!g.transposed && conv_tma_available()) {
        // stride 2, k3, pad 1 as ONE problem: lattice voxel j gathers dy[j + off], off in {0,1}^3 (8 taps), and produces the 8 input
        // voxels 2j + p, p in {0,1}^3, as 8 parity blocks of N.  Per dimension (off, p) -> kernel index k:  (0,0) -> 1, (0,1) -> 2,
        // (1,1) -> 0, (1,0) -> none  (y[o] = sum_k W[k] x[2o + k - 1]); absent combinations are zero blocks of the weight pack.
        const int cp = pad16(n_real);
        const int ld = (g.in_d + 1) / 2, lh = (g.in_h + 1) / 2, lw = (g.in_w + 1) / 2;
        choose_ntile(8 * cp, 1LL * ld * lh * lw, ntile, ntiles);
        ConvProblem P;
        PackDesc K;
        base(P, K);
        P.ntile = ntile; P.ntiles = ntiles; P.n_real = 8 * cp;
        K.ntile = ntile; K.ntiles = ntiles;
        P.istride = 1;
        P.od = ld; P.oh = lh; P.ow = lw;
        P.ostep = 2;
        P.ntaps = 8;
        P.shuffle_cp = cp; P.shuffle_nreal = n_real;
        K.stack_cp = cp;
        const int kof[2][2] = {{1, 2}, {-1, 0}};   // [off][p]
        for (int t = 0; t < 8; ++t) {
            const int oz = t >> 2, oy = (t >> 1) & 1, ox = t & 1;
            P.taps[t] = ConvTap{int8_t(oz), int8_t(oy), int8_t(ox), 0};
            for (int par = 0; par < 8; ++par) {
                const int kz = kof[oz][par >> 2], ky = kof[oy][(par >> 1) & 1], kx = kof[ox][par & 1];
                K.stack_ref[t][par] = (kz < 0 || ky < 0 || kx < 0) ? (signed char)-1 : (signed char)((kz * 3 + ky) * 3 + kx);
            }
        }
        K.dimA = g.cout; K.dimB = g.cin[0] + g.cin[1]; K.ktaps = 27; K.n_is_A = 0; K.ntaps = 8;
        probs.push_back(P);
        packs.push_back(K);
    } else if (!g.transposed) {
        // stride 2, k3, pad 1:  y[o] = sum_k W[k] x[2o + k - 1]  =>  for x index i = 2j + p:
        //   p = 0: k = 1 (o = j);   p = 1: k = 0 (o = j + 1), k = 2 (o = j)
        const int kk[2][2] = {{1, -1}, {0, 2}};
        const int off[2][2] = {{0, 0}, {1, 0}};
        const int cnt[2] = {1, 2};
        for (int pz = 0; pz < 2; ++pz)
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    ConvProblem P;
                    PackDesc K;
                    base(P, K);
                    P.istride = 1;
                    P.od = (g.in_d - pz + 1) / 2; P.oh = (g.in_h - py + 1) / 2; P.ow = (g.in_w - px + 1) / 2;
                    if (P.od <= 0 || P.oh <= 0 || P.ow <= 0) continue;
                    P.ostep = 2; P.ooff_z = pz; P.ooff_y = py; P.ooff_x = px;
                    int t = 0;
                    for (int a = 0; a < cnt[pz]; ++a)
                        for (int b = 0; b < cnt[py]; ++b)
                            for (int c = 0; c < cnt[px]; ++c, ++t) {
                                P.taps[t] = ConvTap{int8_t(off[pz][a]), int8_t(off[py][b]), int8_t(off[px][c]), 0};
                                K.tap_ref[t] = (kk[pz][a] * 3 + kk[py][b]) * 3 + kk[px][c];
                            }
                    P.ntaps = t;
                    K.dimA = g.cout; K.dimB = g.cin[0] + g.cin[1]; K.ktaps = 27; K.n_is_A = 0; K.ntaps = t;
                    probs.push_back(P);
                    packs.push_back(K);
                }
    } else {
        // conv_transpose k2 s2:  y[2i + a] = sum_ci x[i] W[ci][co][a]  =>  dx[i] = sum_a sum_co dy[2i + a] W[ci][co][a]
        ConvProblem P;
        PackDesc K;
        base(P, K);
        P.istride = 2;
        P.ntaps = 8;
        int t = 0;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c, ++t) {
                    P.taps[t] = ConvTap{int8_t(a), int8_t(b), int8_t(c), 0};
                    K.tap_ref[t] = t;
                }
        P.od = g.in_d; P.oh = g.in_h; P.ow = g.in_w;
        P.ostep = 1;
        K.dimA = g.cin[0]; K.dimB = g.cout; K.ktaps = 8; K.n_is_A = 1; K.ntaps = 8;
        probs.push_back(P);
        packs.push_back(K);
    }
}

void plan_wgrad(const LayerGeom& g, int src, WgradProblem& W) {
    std::memset(&W, 0, sizeof(W));
    if (!g.transposed) {
        const int pad = (g.ks - 1) / 2;
        W.t_cp = pad16(g.cin[src]); W.t_coff = 0; W.t_c = pad16(g.cin[src]); W.t_creal = g.cin[src];
        W.t_d = g.in_d; W.t_h = g.in_h; W.t_w = g.in_w;
        W.tstride = g.stride;
        W.ntaps = g.ks * g.ks * g.ks;
        int t = 0;
        for (int kz = 0; kz < g.ks; ++kz)
            for (int ky = 0; ky < g.ks; ++ky)
                for (int kx = 0; kx < g.ks; ++kx, ++t) {
                    W.taps[t] = ConvTap{int8_t(kz - pad), int8_t(ky - pad), int8_t(kx - pad), 0};
                    W.tap_ref[t] = t;
                }
        W.ld = g.out_d; W.lh = g.out_h; W.lw = g.out_w;
        W.u_cp = pad16(g.cout); W.u_coff = 0; W.u_c = pad16(g.cout); W.u_creal = g.cout;
        W.w_mtot = g.cin[0] + g.cin[1]; W.w_moff = src ? g.cin[0] : 0; W.w_ktaps = W.ntaps;
        W.w_ntot = g.cout; W.w_noff = 0;
    } else {
        W.t_cp = pad16(g.cout); W.t_coff = 0; W.t_c = pad16(g.cout); W.t_creal = g.cout;
        W.t_d = g.out_d; W.t_h = g.out_h; W.t_w = g.out_w;
        W.tstride = 2;
        W.ntaps = 8;
        int t = 0;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c, ++t) {
                    W.taps[t] = ConvTap{int8_t(a), int8_t(b), int8_t(c), 0};
                    W.tap_ref[t] = t;
                }
        W.ld = g.in_d; W.lh = g.in_h; W.lw = g.in_w;
        W.u_cp = pad16(g.cin[0]); W.u_coff = 0; W.u_c = pad16(g.cin[0]); W.u_creal = g.cin[0];
        W.w_mtot = g.cout; W.w_moff = 0; W.w_ktaps = 8;
        W.w_ntot = g.cin[0]; W.w_noff = 0;
    }
}

}  // namespace u3d
