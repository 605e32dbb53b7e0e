// K7 host generators: Type-I single-panel codebook tables (TS 38.214 Tables 5.2.2.2.1-1 ... -12, as coded
// at dlPMISelect.m:853-1349 and pmiType1SinglePanelCodebook.m:46-554) and PUSCH codebooks
// (TS 38.211 Tables 6.3.1.5-1 ... -7; the reference calls the toolbox's nrPUSCHCodebook, pmiSelect.m:45).
#include "codebook.cuh"
#include <cmath>

namespace isac {

using cd = std::complex<double>;

void subband_info(bool subbandMode, int nStartBWP, int nSizeBWP, int nsbprb, std::vector<int>& sizes) {
    sizes.clear();
    if (!subbandMode || nSizeBWP < 24 || nsbprb <= 0) {
        sizes.push_back(nSizeBWP);
        return;
    }
    const int first = nsbprb - (nStartBWP % nsbprb);
    int last = (nStartBWP + nSizeBWP) % nsbprb;
    if (last == 0) last = nsbprb;
    const int n = (nSizeBWP - (first + last)) / nsbprb + 2;
    sizes.assign(n, nsbprb);
    sizes[0] = first;
    sizes[n - 1] = last;
}

namespace {

struct Builder {
    const CsiConfig& c;
    CodebookTable& t;
    int N1, N2, O1, O2, P;
    bool restrictedLM(std::initializer_list<int> bits) const {
        if (!c.subsetRestriction) return false;
        const int n = N1 * O1 * N2 * O2;  // bits beyond the vector never match (isRestricted, dlPMISelect.m:1811-1814)
        for (int b : bits)
            if (b >= 0 && b < n && !c.subsetRestriction[b]) return true;
        return false;
    }
    bool restrictedI2(int n) const { return c.i2Restriction && n >= 0 && n < 16 && !c.i2Restriction[n]; }
    // beam index of v_{l,m} (periodic in l, m)
    int beam(int l, int m) const {
        const int L1 = N1 * O1, L2 = N2 * O2;
        return ((l % L1 + L1) % L1) * L2 + ((m % L2 + L2) % L2);
    }
    size_t cand(int i2, int i11, int i12, int i13) const {
        return (size_t)i2 + (size_t)t.n2 * (i11 + (size_t)t.n11 * (i12 + (size_t)t.n12 * i13));
    }
    void alloc(int n2, int n11, int n12, int n13) {
        t.n2 = n2; t.n11 = n11; t.n12 = n12; t.n13 = n13;
        t.valid.assign(t.nCand(), 0);
        t.layers.assign((size_t)t.nCand() * t.nLayers, LayerDesc{});
    }
    // set column j of candidate `ci` to [c0*v_b ; c1*v_b (; c2*v_b ; c3*v_b)]
    void col(size_t ci, int j, int b, cd c0, cd c1, cd c2 = 0.0, cd c3 = 0.0) {
        LayerDesc& d = t.layers[ci * t.nLayers + j];
        d.beam = b;
        d.coef[0] = c0; d.coef[1] = c1; d.coef[2] = c2; d.coef[3] = c3;
        for (int b = 4; b < kMaxBlocks; ++b) d.coef[b] = 0.0;
    }
};

cd phi(int n) {  // exp(1i*pi*n/2), exact
    static const cd tab[4] = {cd(1, 0), cd(0, 1), cd(-1, 0), cd(0, -1)};
    return tab[((n % 4) + 4) % 4];
}

}  // namespace

int build_type1sp_table(Ctx* ctx, const CsiConfig& c, int nu, int variant, CodebookTable& t) {
    t = CodebookTable();
    const int P = c.nPorts;
    if (nu < 1 || nu > kMaxLayers || nu > P) {
        set_error(ctx, "nr5g:hDLPMISelect:InvalidNumLayers");
        return kErrInvalidArg;
    }
    t.P = P;
    t.nLayers = nu;
    Builder B{c, t, c.N1, c.N2, c.O1, c.O2, P};
    const int N1 = c.N1, N2 = c.N2, O1 = c.O1, O2 = c.O2;
    if (P == 1) {  // W = 1 (dlPMISelect.m:328-331)
        t.NB = 1; t.Pb = 1; t.nBeams = 1; t.beams = {cd(1, 0)}; t.scale = 1.0;
        B.alloc(1, 1, 1, 1);
        t.valid[0] = 1;
        B.col(0, 0, 0, 1.0, 0.0);
        return kOk;
    }
    if (P == 2) {  // Table 5.2.2.2.1-1 (dlPMISelect.m:891-916)
        if (nu > 2) { set_error(ctx, "2 ports support at most 2 layers"); return kErrInvalidArg; }
        t.NB = 2; t.Pb = 1; t.nBeams = 1; t.beams = {cd(1, 0)};
        if (nu == 1) {
            t.scale = 1.0 / std::sqrt(2.0);
            B.alloc(4, 1, 1, 1);
            for (int i = 0; i < 4; ++i) {
                t.valid[i] = !(c.subsetRestriction && !c.subsetRestriction[i]);
                B.col(i, 0, 0, 1.0, phi(i));
            }
        } else {
            t.scale = 0.5;
            B.alloc(2, 1, 1, 1);
            for (int i = 0; i < 2; ++i) {
                t.valid[i] = !(c.subsetRestriction && !c.subsetRestriction[4 + i]);
                B.col(i, 0, 0, 1.0, phi(i));
                B.col(i, 1, 0, 1.0, -phi(i));
            }
        }
        return kOk;
    }
    if (2 * N1 * N2 != P) { set_error(ctx, "nr5g:dlPMISelect:InvalidPanelDimensions"); return kErrInvalidArg; }
    const bool vbar = (nu == 3 || nu == 4) && P >= 16;
    t.scale = 1.0 / std::sqrt((double)nu * P);
    if (!vbar) {
        t.NB = 2; t.Pb = P / 2; t.nBeams = N1 * O1 * N2 * O2;
        t.beams.resize((size_t)t.nBeams * t.Pb);
        for (int l = 0; l < N1 * O1; ++l)
            for (int m = 0; m < N2 * O2; ++m)
                for (int n1 = 0; n1 < N1; ++n1)
                    for (int n2 = 0; n2 < N2; ++n2) {  // getVlm: reshape((ul.*um).',[],1) -> N2 fastest
                        const double ang = 2.0 * M_PI * ((double)l * n1 / (O1 * N1) + (double)m * n2 / (O2 * N2));
                        t.beams[(size_t)(l * N2 * O2 + m) * t.Pb + n1 * N2 + n2] = cd(std::cos(ang), std::sin(ang));
                    }
    } else {
        t.NB = 4; t.Pb = P / 4; t.nBeams = (N1 * O1 / 2) * N2 * O2;
        t.beams.resize((size_t)t.nBeams * t.Pb);
        for (int l = 0; l < N1 * O1 / 2; ++l)
            for (int m = 0; m < N2 * O2; ++m)
                for (int n1 = 0; n1 < N1 / 2; ++n1)
                    for (int n2 = 0; n2 < N2; ++n2) {  // getVbarlm
                        const double ang = 2.0 * M_PI * ((double)l * n1 / (O1 * N1 / 2.0) + (double)m * n2 / (O2 * N2));
                        t.beams[(size_t)(l * N2 * O2 + m) * t.Pb + n1 * N2 + n2] = cd(std::cos(ang), std::sin(ang));
                    }
    }
    static const int lmAdd[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
    const int mode = c.codebookMode;
    if (nu == 1) {
        if (mode == 1) {
            B.alloc(4, N1 * O1, N2 * O2, 1);
            for (int i11 = 0; i11 < t.n11; ++i11)
                for (int i12 = 0; i12 < t.n12; ++i12)
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const size_t ci = B.cand(i2, i11, i12, 0);
                        if (B.restrictedLM({N2 * O2 * i11 + i12}) || B.restrictedI2(i2)) continue;
                        t.valid[ci] = 1;
                        B.col(ci, 0, B.beam(i11, i12), 1.0, phi(i2));
                    }
        } else {
            B.alloc(16, N1 * O1 / 2, N2 == 1 ? 1 : N2 * O2 / 2, 1);
            for (int i11 = 0; i11 < t.n11; ++i11)
                for (int i12 = 0; i12 < t.n12; ++i12)
                    for (int i2 = 0; i2 < 16; ++i2) {
                        const int f = i2 / 4;
                        const int l = N2 == 1 ? 2 * i11 + f : 2 * i11 + lmAdd[f][0];
                        const int m = N2 == 1 ? 0 : 2 * i12 + lmAdd[f][1];
                        const size_t ci = B.cand(i2, i11, i12, 0);
                        if (B.restrictedLM({N2 * O2 * l + m}) || B.restrictedI2(i2)) continue;
                        t.valid[ci] = 1;
                        B.col(ci, 0, B.beam(l, m), 1.0, phi(i2 % 4));
                    }
        }
        return kOk;
    }
    if (nu == 2) {
        std::vector<int> k1, k2;  // Table 5.2.2.2.1-3
        if (N1 > N2 && N2 > 1) { k1 = {0, O1, 0, 2 * O1}; k2 = {0, 0, O2, 0}; }
        else if (N1 == N2) { k1 = {0, O1, 0, O1}; k2 = {0, 0, O2, O2}; }
        else if (N1 == 2 && N2 == 1) { k1 = {0, O1}; k2 = {0, 0}; }
        else { k1 = {0, O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0, 0}; }
        const int n13 = (int)k1.size();
        if (mode == 1) {
            B.alloc(2, N1 * O1, N2 * O2, n13);
            for (int i11 = 0; i11 < t.n11; ++i11)
                for (int i12 = 0; i12 < t.n12; ++i12)
                    for (int i13 = 0; i13 < n13; ++i13)
                        for (int i2 = 0; i2 < 2; ++i2) {
                            const size_t ci = B.cand(i2, i11, i12, i13);
                            if (B.restrictedLM({N2 * O2 * i11 + i12}) || B.restrictedI2(i2)) continue;
                            t.valid[ci] = 1;
                            B.col(ci, 0, B.beam(i11, i12), 1.0, phi(i2));
                            B.col(ci, 1, B.beam(i11 + k1[i13], i12 + k2[i13]), 1.0, -phi(i2));
                        }
        } else {
            B.alloc(8, N1 * O1 / 2, N2 == 1 ? 1 : N2 * O2 / 2, n13);
            for (int i11 = 0; i11 < t.n11; ++i11)
                for (int i12 = 0; i12 < t.n12; ++i12)
                    for (int i13 = 0; i13 < n13; ++i13)
                        for (int i2 = 0; i2 < 8; ++i2) {
                            const int f = i2 / 2;
                            const int fp = variant == kVariantGNB ? i2 / 4 : f;  // pmiType1SinglePanelCodebook.m:225,227
                            int l, lp, m, mp;
                            if (N2 == 1) { l = 2 * i11 + f; lp = 2 * i11 + f + k1[i13]; m = 0; mp = 0; }
                            else {
                                l = 2 * i11 + lmAdd[f][0]; lp = 2 * i11 + k1[i13] + lmAdd[fp][0];
                                m = 2 * i12 + lmAdd[f][1]; mp = 2 * i12 + k2[i13] + lmAdd[fp][1];
                            }
                            const size_t ci = B.cand(i2, i11, i12, i13);
                            if (B.restrictedLM({N2 * O2 * l + m}) || B.restrictedI2(i2)) continue;
                            t.valid[ci] = 1;
                            B.col(ci, 0, B.beam(l, m), 1.0, phi(i2 % 2));
                            B.col(ci, 1, B.beam(lp, mp), 1.0, -phi(i2 % 2));
                        }
        }
        return kOk;
    }
    if (nu == 3 || nu == 4) {
        if (!vbar) {
            std::vector<int> k1, k2;  // Table 5.2.2.2.1-4
            if (N1 == 2 && N2 == 1) { k1 = {O1}; k2 = {0}; }
            else if (N1 == 4 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0}; }
            else if (N1 == 6 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1, 4 * O1}; k2 = {0, 0, 0, 0}; }
            else if (N1 == 2 && N2 == 2) { k1 = {O1, 0, O1}; k2 = {0, O2, O2}; }
            else if (N1 == 3 && N2 == 2) { k1 = {O1, 0, O1, 2 * O1}; k2 = {0, O2, O2, 0}; }
            else { set_error(ctx, "unsupported panel for 3-4 layers"); return kErrUnsupported; }
            const int n13 = (int)k1.size();
            B.alloc(2, N1 * O1, N2 * O2, n13);
            for (int i11 = 0; i11 < t.n11; ++i11)
                for (int i12 = 0; i12 < t.n12; ++i12)
                    for (int i13 = 0; i13 < n13; ++i13)
                        for (int i2 = 0; i2 < 2; ++i2) {
                            const size_t ci = B.cand(i2, i11, i12, i13);
                            if (B.restrictedLM({N2 * O2 * i11 + i12}) || B.restrictedI2(i2)) continue;
                            t.valid[ci] = 1;
                            const int b0 = B.beam(i11, i12), b1 = B.beam(i11 + k1[i13], i12 + k2[i13]);
                            const cd ph = phi(i2);
                            B.col(ci, 0, b0, 1.0, ph);
                            B.col(ci, 1, b1, 1.0, ph);
                            B.col(ci, 2, b0, 1.0, -ph);
                            if (nu == 4) B.col(ci, 3, b1, 1.0, -ph);
                        }
            return kOk;
        }
        B.alloc(2, N1 * O1 / 2, N2 * O2, 4);
        const int L12 = N1 * O1 * N2 * O2;
        for (int i11 = 0; i11 < t.n11; ++i11)
            for (int i12 = 0; i12 < t.n12; ++i12)
                for (int i13 = 0; i13 < 4; ++i13)
                    for (int i2 = 0; i2 < 2; ++i2) {
                        const int l = i11, m = i12;
                        const int b0 = ((N2 * O2 * (2 * l - 1) + m) % L12 + L12) % L12;
                        if (B.restrictedLM({b0, N2 * O2 * (2 * l) + m, N2 * O2 * (2 * l + 1) + m}) || B.restrictedI2(i2)) continue;
                        // gNB copy writes W(:,:,i2+1,i11+1,i12+1) for every i13: slice 1 ends up holding i13 = 3
                        // and slices 2..4 stay zero (pmiType1SinglePanelCodebook.m:348,358)
                        const size_t ci = B.cand(i2, i11, i12, variant == kVariantGNB ? 0 : i13);
                        t.valid[ci] = 1;
                        const double a = M_PI * i13 / 4.0;
                        const cd th(std::cos(a), std::sin(a)), ph = phi(i2);
                        const int b = l * N2 * O2 + m;
                        B.col(ci, 0, b, 1.0, th, ph, ph * th);
                        B.col(ci, 1, b, 1.0, -th, ph, -ph * th);
                        B.col(ci, 2, b, 1.0, th, -ph, -ph * th);
                        if (nu == 4) B.col(ci, 3, b, 1.0, -th, -ph, ph * th);
                    }
        return kOk;
    }
    if (nu == 5 || nu == 6) {
        B.alloc(2, N1 * O1, N2 == 1 ? 1 : N2 * O2, 1);
        for (int i11 = 0; i11 < t.n11; ++i11)
            for (int i12 = 0; i12 < t.n12; ++i12)
                for (int i2 = 0; i2 < 2; ++i2) {
                    int l = i11, lp = i11 + O1, ld, m, mp, md;
                    if (N2 == 1) { ld = i11 + 2 * O1; m = mp = md = 0; }
                    else { ld = i11 + O1; m = i12; mp = i12; md = i12 + O2; }
                    const size_t ci = B.cand(i2, i11, i12, 0);
                    if (B.restrictedLM({N2 * O2 * l + m}) || B.restrictedI2(i2)) continue;
                    t.valid[ci] = 1;
                    const int b = B.beam(l, m), bp = B.beam(lp, mp), bd = B.beam(ld, md);
                    const cd ph = phi(i2);
                    B.col(ci, 0, b, 1.0, ph);
                    B.col(ci, 1, b, 1.0, -ph);
                    if (nu == 5) {
                        B.col(ci, 2, bp, 1.0, 1.0);
                        B.col(ci, 3, bp, 1.0, -1.0);
                        B.col(ci, 4, bd, 1.0, 1.0);
                    } else {
                        B.col(ci, 2, bp, 1.0, ph);
                        B.col(ci, 3, bp, 1.0, -ph);
                        B.col(ci, 4, bd, 1.0, 1.0);
                        B.col(ci, 5, bd, 1.0, -1.0);
                    }
                }
        return kOk;
    }
    // 7 or 8 layers
    int n11, n12;
    if (N2 == 1) { n12 = 1; n11 = (N1 == 4) ? N1 * O1 / 2 : N1 * O1; }
    else { n11 = N1 * O1; n12 = ((N1 == 2 && N2 == 2) || (N1 > 2 && N2 > 2)) ? N2 * O2 : N2 * O2 / 2; }
    B.alloc(2, n11, n12, 1);
    for (int i11 = 0; i11 < n11; ++i11)
        for (int i12 = 0; i12 < n12; ++i12)
            for (int i2 = 0; i2 < 2; ++i2) {
                int ls[4], ms[4];
                if (N2 == 1) { for (int q = 0; q < 4; ++q) { ls[q] = i11 + q * O1; ms[q] = 0; } }
                else { ls[0] = i11; ls[1] = i11 + O1; ls[2] = i11; ls[3] = i11 + O1; ms[0] = i12; ms[1] = i12; ms[2] = i12 + O2; ms[3] = i12 + O2; }
                const size_t ci = B.cand(i2, i11, i12, 0);
                if (B.restrictedLM({N2 * O2 * ls[0] + ms[0]}) || B.restrictedI2(i2)) continue;
                t.valid[ci] = 1;
                int b[4];
                for (int q = 0; q < 4; ++q) b[q] = B.beam(ls[q], ms[q]);
                const cd ph = phi(i2);
                if (nu == 7) {
                    B.col(ci, 0, b[0], 1.0, ph);  B.col(ci, 1, b[0], 1.0, -ph); B.col(ci, 2, b[1], 1.0, ph);
                    B.col(ci, 3, b[2], 1.0, 1.0); B.col(ci, 4, b[2], 1.0, -1.0);
                    B.col(ci, 5, b[3], 1.0, 1.0); B.col(ci, 6, b[3], 1.0, -1.0);
                } else {
                    B.col(ci, 0, b[0], 1.0, ph);  B.col(ci, 1, b[0], 1.0, -ph);
                    B.col(ci, 2, b[1], 1.0, ph);  B.col(ci, 3, b[1], 1.0, -ph);
                    B.col(ci, 4, b[2], 1.0, 1.0); B.col(ci, 5, b[2], 1.0, -1.0);
                    B.col(ci, 6, b[3], 1.0, 1.0); B.col(ci, 7, b[3], 1.0, -1.0);
                }
            }
    return kOk;
}

// ------------------------------------------------------------------------------------------
// PUSCH codebooks, TS 38.211 Tables 6.3.1.5-1 ... -7 (transform precoding disabled)
// ------------------------------------------------------------------------------------------
namespace {
const cd J(0, 1);
struct Mat { int rows, cols; std::vector<cd> v; double div; };  // row-major entries, W = v / div

std::vector<Mat> pusch_mats(int nu, int P) {
    std::vector<Mat> out;
    auto add = [&](int r, int c, std::vector<cd> v, double d) { out.push_back(Mat{r, c, std::move(v), d}); };
    const double s2 = std::sqrt(2.0), s3 = std::sqrt(3.0);
    if (P == 1) { add(1, 1, {1}, 1); return out; }
    if (P == 2 && nu == 1) {
        const cd t[6][2] = {{1, 0}, {0, 1}, {1, 1}, {1, -1}, {1, J}, {1, -J}};
        for (auto& r : t) add(2, 1, {r[0], r[1]}, s2);
        return out;
    }
    if (P == 2 && nu == 2) {
        add(2, 2, {1, 0, 0, 1}, s2); add(2, 2, {1, 1, 1, -1}, 2); add(2, 2, {1, 1, J, -J}, 2);
        return out;
    }
    if (nu == 1) {
        const cd t[28][4] = {{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1},{1,0,1,0},{1,0,-1,0},{1,0,J,0},{1,0,-J,0},
            {0,1,0,1},{0,1,0,-1},{0,1,0,J},{0,1,0,-J},{1,1,1,1},{1,1,J,J},{1,1,-1,-1},{1,1,-J,-J},
            {1,J,1,J},{1,J,J,-1},{1,J,-1,-J},{1,J,-J,1},{1,-1,1,-1},{1,-1,J,-J},{1,-1,-1,1},{1,-1,-J,J},
            {1,-J,1,-J},{1,-J,J,1},{1,-J,-1,J},{1,-J,-J,-1}};
        for (auto& r : t) add(4, 1, {r[0], r[1], r[2], r[3]}, 2);
        return out;
    }
    if (nu == 2) {
        const int sel[6][2] = {{0,1},{0,2},{0,3},{1,2},{1,3},{2,3}};
        for (auto& s : sel) { std::vector<cd> v(8, 0.0); v[s[0] * 2 + 0] = 1; v[s[1] * 2 + 1] = 1; add(4, 2, v, 2); }
        const cd ab[8][2] = {{1,-J},{1,J},{-J,1},{-J,-1},{-1,-J},{-1,J},{J,1},{J,-1}};
        for (auto& p : ab) add(4, 2, {1, 0, 0, 1, p[0], 0, 0, p[1]}, 2);
        const cd fc[8][8] = {{1,1,1,1,1,-1,1,-1},{1,1,1,1,J,-J,J,-J},{1,1,J,J,1,-1,J,-J},{1,1,J,J,J,-J,-1,1},
            {1,1,-1,-1,1,-1,-1,1},{1,1,-1,-1,J,-J,-J,J},{1,1,-J,-J,1,-1,-J,J},{1,1,-J,-J,J,-J,1,-1}};
        for (auto& r : fc) add(4, 2, std::vector<cd>(r, r + 8), 2 * s2);
        return out;
    }
    if (nu == 3) {
        add(4, 3, {1,0,0, 0,1,0, 0,0,1, 0,0,0}, 2);
        add(4, 3, {1,0,0, 0,1,0, 1,0,0, 0,0,1}, 2);
        add(4, 3, {1,0,0, 0,1,0, -1,0,0, 0,0,1}, 2);
        add(4, 3, {1,1,1, 1,-1,1, 1,1,-1, 1,-1,-1}, 2 * s3);
        add(4, 3, {1,1,1, 1,-1,1, J,J,-J, J,-J,-J}, 2 * s3);
        add(4, 3, {1,1,1, -1,1,-1, 1,1,-1, -1,1,1}, 2 * s3);
        add(4, 3, {1,1,1, -1,1,-1, J,J,-J, -J,J,J}, 2 * s3);
        return out;
    }
    add(4, 4, {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, 2);
    add(4, 4, {1,1,0,0, 0,0,1,1, 1,-1,0,0, 0,0,1,-1}, 2 * s2);
    add(4, 4, {1,1,0,0, 0,0,1,1, J,-J,0,0, 0,0,J,-J}, 2 * s2);
    add(4, 4, {1,1,1,1, 1,-1,1,-1, 1,1,-1,-1, 1,-1,-1,1}, 4);
    add(4, 4, {1,1,1,1, 1,-1,1,-1, J,J,-J,-J, J,-J,-J,J}, 4);
    return out;
}
}  // namespace

int build_pusch_table(Ctx* ctx, int nu, int P, CodebookTable& t) {
    t = CodebookTable();
    if (!(P == 1 || P == 2 || P == 4)) {
        set_error(ctx, "nr5g:hMaxPUSCHPrecodingMatrixIndicator:InvalidNPorts");
        return kErrInvalidArg;
    }
    if (nu < 1 || nu > P) {
        set_error(ctx, "nr5g:hMaxPUSCHPrecodingMatrixIndicator:TooManyLayers");
        return kErrInvalidArg;
    }
    std::vector<Mat> mats = pusch_mats(nu, P);
    t.P = P; t.nLayers = nu; t.NB = P; t.Pb = 1; t.nBeams = 1; t.beams = {cd(1, 0)}; t.scale = 1.0;
    t.n2 = (int)mats.size(); t.n11 = t.n12 = t.n13 = 1;
    t.valid.assign(t.n2, 1);
    t.layers.assign((size_t)t.n2 * nu, LayerDesc{});
    t.candScale.resize(t.n2);
    for (int c = 0; c < t.n2; ++c) {
        t.candScale[c] = 1.0 / mats[c].div;
        for (int j = 0; j < nu; ++j) {
            LayerDesc& d = t.layers[(size_t)c * nu + j];
            d.beam = 0;
            for (int p = 0; p < P; ++p) d.coef[p] = mats[c].v[p * nu + j];
        }
    }
    return kOk;
}

// getPMIType1MultiPanelCodebook (dlPMISelect.m:1351-1772; TS 38.214 Tables 5.2.2.2.2-1..-6) as the explicit array the
// reference returns: W [P x nu x i20 x i21 x i22 x i11 x i12 x i13 x i141 x i142 x i143], restricted precoders all zero.
// Every column is a "+" column  [c_g v ;  c_g phi_n v]_g  or a "-" column  [c_g v ; -c_g phi_n v]_g  over the Ng panels g,
// with c_0 = 1 and c_g = phi(i14g) in codebook mode 1; mode 2 (Ng = 2): panel 1 carries a(i141) b(i21) v on the first and
// +-a(i142) b(i22) v on the second polarisation.  Layers: [+v], [+v, -v'], [+v, +v', -v], [+v, +v', -v, -v'] with
// v' = v_{l+k1, m+k2} (k tables :1518-1536 for two layers, :1615-1636 for three and four).  Pure host code.
int type1mp_codebook(Ctx* ctx, const CsiConfig& c, int Ng, int nu, int dims[9], std::vector<cd>* W) {
    const int N1 = c.N1, N2 = c.N2, O1 = c.O1, O2 = c.O2, mode = c.codebookMode;
    if ((Ng != 2 && Ng != 4) || (mode != 1 && mode != 2) || (mode == 2 && Ng != 2) || nu < 1 || nu > 4 || N1 < 1 || N2 < 1) {
        set_error(ctx, "nr5g:dlPMISelect:InvalidPanelDimensions");
        return kErrInvalidArg;
    }
    const int P = 2 * Ng * N1 * N2, Pb = N1 * N2;
    const int n11 = N1 * O1, n12 = N2 * O2, n141 = 4;
    const int n142 = mode == 1 ? (Ng == 2 ? 1 : 4) : 4, n143 = mode == 1 ? (Ng == 2 ? 1 : 4) : 1;
    const int n21 = mode == 1 ? 1 : 2, n22 = n21;
    std::vector<int> k1, k2;
    int n20 = 2;
    if (nu == 1) { n20 = 4; k1 = {0}; k2 = {0}; }
    else if (nu == 2) {
        if (N1 > N2 && N2 > 1) { k1 = {0, O1, 0, 2 * O1}; k2 = {0, 0, O2, 0}; }
        else if (N1 == N2) { k1 = {0, O1, 0, O1}; k2 = {0, 0, O2, O2}; }
        else if (N1 == 2 && N2 == 1) { k1 = {0, O1}; k2 = {0, 0}; }
        else { k1 = {0, O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0, 0}; }
    } else {
        if (N1 == 2 && N2 == 1) { k1 = {O1}; k2 = {0}; }
        else if (N1 == 4 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0}; }
        else if (N1 == 8 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1, 4 * O1}; k2 = {0, 0, 0, 0}; }
        else if (N1 == 2 && N2 == 2) { k1 = {O1, 0, O1}; k2 = {0, O2, O2}; }
        else if (N1 == 4 && N2 == 2) { k1 = {O1, 0, O1, 2 * O1}; k2 = {0, O2, O2, 0}; }
        else { set_error(ctx, "nr5g:dlPMISelect:InvalidPanelDimensions"); return kErrInvalidArg; }
    }
    const int n13 = (int)k1.size();
    const int d[9] = {n20, n21, n22, n11, n12, n13, n141, n142, n143};
    for (int i = 0; i < 9; ++i) dims[i] = d[i];
    if (!W) return kOk;
    size_t nCand = 1;
    for (int i = 0; i < 9; ++i) nCand *= (size_t)d[i];
    W->assign((size_t)P * nu * nCand, cd(0, 0));
    const double s = 1.0 / std::sqrt((double)nu * P);
    const cd A0(std::sqrt(0.5), std::sqrt(0.5)), B0(std::sqrt(0.5), -std::sqrt(0.5));   // a(x) = A0 phi(x), b(x) = B0 phi(x)
    auto vlm = [&](int l, int m, std::vector<cd>& v) {   // getVlm: N2 fastest
        v.resize(Pb);
        for (int a1 = 0; a1 < N1; ++a1)
            for (int a2 = 0; a2 < N2; ++a2) {
                const double ang = 2.0 * M_PI * ((double)l * a1 / (O1 * N1) + (double)m * a2 / (O2 * N2));
                v[a1 * N2 + a2] = cd(std::cos(ang), std::sin(ang));
            }
    };
    std::vector<cd> v, vp;
    size_t ci = 0;   // MATLAB linear order over [i20 i21 i22 i11 i12 i13 i141 i142 i143]: walk it with nested loops, last index outermost
    for (int i143 = 0; i143 < n143; ++i143)
     for (int i142 = 0; i142 < n142; ++i142)
      for (int i141 = 0; i141 < n141; ++i141)
       for (int i13 = 0; i13 < n13; ++i13)
        for (int i12 = 0; i12 < n12; ++i12)
         for (int i11 = 0; i11 < n11; ++i11) {
            const int bit = N2 * O2 * i11 + i12;
            const bool restricted = c.subsetRestriction && bit < n11 * n12 && !c.subsetRestriction[bit];
            if (!restricted) { vlm(i11, i12, v); vlm(i11 + k1[i13], i12 + k2[i13], vp); }
            for (int i22 = 0; i22 < n22; ++i22)
             for (int i21 = 0; i21 < n21; ++i21)
              for (int i20 = 0; i20 < n20; ++i20, ++ci) {
                if (restricted) continue;
                const cd fn = phi(i20);
                cd first[4], second[4];   // per panel: coefficient of the first / second polarisation block of a "+" column
                if (mode == 1) {
                    const cd cg[4] = {cd(1, 0), phi(i141), phi(i142), phi(i143)};
                    for (int g = 0; g < Ng; ++g) { first[g] = cg[g]; second[g] = cg[g] * fn; }
                } else {
                    first[0] = cd(1, 0); second[0] = fn;
                    first[1] = A0 * phi(i141) * B0 * phi(i21); second[1] = A0 * phi(i142) * B0 * phi(i22);
                }
                for (int j = 0; j < nu; ++j) {
                    const bool prime = (nu == 2 && j == 1) || (nu >= 3 && (j == 1 || j == 3));
                    const bool minus = (nu == 2 && j == 1) || (nu >= 3 && j >= 2);
                    const std::vector<cd>& x = prime ? vp : v;
                    cd* col = W->data() + ((size_t)ci * nu + j) * P;
                    for (int g = 0; g < Ng; ++g)
                        for (int q = 0; q < Pb; ++q) {
                            col[(2 * g) * Pb + q] = s * first[g] * x[q];
                            col[(2 * g + 1) * Pb + q] = (minus ? -s : s) * second[g] * x[q];
                        }
                }
              }
         }
    return kOk;
}

int build_type1mp_table(Ctx* ctx, const CsiConfig& c, int nu, CodebookTable& t) {
    t = CodebookTable();
    const int Ng = c.nPanels, N1 = c.N1, N2 = c.N2, O1 = c.O1, O2 = c.O2, mode = c.codebookMode;
    int d[9];
    int st = type1mp_codebook(ctx, c, Ng, nu, d, nullptr);   // validates the configuration, returns the index-set lengths
    if (st) return st;
    if (2 * Ng * N1 * N2 != c.nPorts) { set_error(ctx, "nr5g:dlPMISelect:InvalidPanelDimensions"); return kErrInvalidArg; }
    const int n20 = d[0], n21 = d[1], n22 = d[2], n11 = d[3], n12 = d[4], n13 = d[5], n141 = d[6], n142 = d[7], n143 = d[8];
    std::vector<int> k1, k2;   // same tables as type1mp_codebook
    if (nu == 1) { k1 = {0}; k2 = {0}; }
    else if (nu == 2) {
        if (N1 > N2 && N2 > 1) { k1 = {0, O1, 0, 2 * O1}; k2 = {0, 0, O2, 0}; }
        else if (N1 == N2) { k1 = {0, O1, 0, O1}; k2 = {0, 0, O2, O2}; }
        else if (N1 == 2 && N2 == 1) { k1 = {0, O1}; k2 = {0, 0}; }
        else { k1 = {0, O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0, 0}; }
    } else {
        if (N1 == 2 && N2 == 1) { k1 = {O1}; k2 = {0}; }
        else if (N1 == 4 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1}; k2 = {0, 0, 0}; }
        else if (N1 == 8 && N2 == 1) { k1 = {O1, 2 * O1, 3 * O1, 4 * O1}; k2 = {0, 0, 0, 0}; }
        else if (N1 == 2 && N2 == 2) { k1 = {O1, 0, O1}; k2 = {0, O2, O2}; }
        else { k1 = {O1, 0, O1, 2 * O1}; k2 = {0, O2, O2, 0}; }
    }
    t.P = c.nPorts;
    t.nLayers = nu;
    t.NB = 2 * Ng;
    t.Pb = N1 * N2;
    t.nBeams = n11 * n12;
    t.scale = 1.0 / std::sqrt((double)nu * t.P);
    t.beams.resize((size_t)t.nBeams * t.Pb);
    for (int l = 0; l < n11; ++l)
        for (int m = 0; m < n12; ++m)
            for (int a1 = 0; a1 < N1; ++a1)
                for (int a2 = 0; a2 < N2; ++a2) {   // getVlm: N2 fastest
                    const double ang = 2.0 * M_PI * ((double)l * a1 / (O1 * N1) + (double)m * a2 / (O2 * N2));
                    t.beams[(size_t)(l * n12 + m) * t.Pb + a1 * N2 + a2] = cd(std::cos(ang), std::sin(ang));
                }
    Builder B{c, t, N1, N2, O1, O2, t.P};
    B.alloc(n20 * n21 * n22, n11, n12, n13 * n141 * n142 * n143);
    const int mp[7] = {n20, n21, n22, n13, n141, n142, n143};
    for (int i = 0; i < 7; ++i) t.mp[i] = mp[i];
    const cd A0(std::sqrt(0.5), std::sqrt(0.5)), B0(std::sqrt(0.5), -std::sqrt(0.5));
    for (int i143 = 0; i143 < n143; ++i143)
     for (int i142 = 0; i142 < n142; ++i142)
      for (int i141 = 0; i141 < n141; ++i141)
       for (int i13 = 0; i13 < n13; ++i13)
        for (int i12 = 0; i12 < n12; ++i12)
         for (int i11 = 0; i11 < n11; ++i11) {
            if (B.restrictedLM({N2 * O2 * i11 + i12})) continue;   // only the v_lm restriction applies (:1434)
            const int bv = B.beam(i11, i12), bvp = B.beam(i11 + k1[i13], i12 + k2[i13]);
            const int i13f = i13 + n13 * (i141 + n141 * (i142 + n142 * i143));
            for (int i22 = 0; i22 < n22; ++i22)
             for (int i21 = 0; i21 < n21; ++i21)
              for (int i20 = 0; i20 < n20; ++i20) {
                const size_t ci = B.cand(i20 + n20 * (i21 + n21 * i22), i11, i12, i13f);
                t.valid[ci] = 1;
                const cd fn = phi(i20);
                cd first[4], second[4];
                if (mode == 1) {
                    const cd cg[4] = {cd(1, 0), phi(i141), phi(i142), phi(i143)};
                    for (int g = 0; g < Ng; ++g) { first[g] = cg[g]; second[g] = cg[g] * fn; }
                } else {
                    first[0] = cd(1, 0); second[0] = fn;
                    first[1] = A0 * phi(i141) * B0 * phi(i21); second[1] = A0 * phi(i142) * B0 * phi(i22);
                }
                for (int j = 0; j < nu; ++j) {
                    const bool prime = (nu == 2 && j == 1) || (nu >= 3 && (j == 1 || j == 3));
                    const bool minus = (nu == 2 && j == 1) || (nu >= 3 && j >= 2);
                    LayerDesc& ld = t.layers[ci * nu + j];
                    ld.beam = prime ? bvp : bv;
                    for (int g = 0; g < Ng; ++g) {
                        ld.coef[2 * g] = first[g];
                        ld.coef[2 * g + 1] = minus ? -second[g] : second[g];
                    }
                }
              }
         }
    return kOk;
}

void materialize_codebook(const CodebookTable& t, std::vector<cd>& W) {
    const int nc = t.nCand();
    W.assign((size_t)t.P * t.nLayers * nc, cd(0, 0));
    for (int c = 0; c < nc; ++c) {
        if (!t.valid[c]) continue;
        const double sc = t.candScale.empty() ? t.scale : t.candScale[c];
        for (int j = 0; j < t.nLayers; ++j) {
            const LayerDesc& d = t.layers[(size_t)c * t.nLayers + j];
            for (int b = 0; b < t.NB; ++b)
                for (int q = 0; q < t.Pb; ++q)
                    W[((size_t)c * t.nLayers + j) * t.P + b * t.Pb + q] = sc * d.coef[b] * t.beams[(size_t)d.beam * t.Pb + q];
        }
    }
}

}  // namespace isac
