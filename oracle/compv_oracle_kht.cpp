/*
 * TEST INFRASTRUCTURE ONLY (see compv_oracle.c).  CPU restatement of the kernel-based Hough transform (KHT) line detector.
 * Reference: core/features/hough/compv_core_feature_houghkht.cxx (all line numbers below), KHT_TYP = double
 * (core/include/compv/core/features/hough/compv_core_feature_houghkht.h:22).
 *   rho/theta tables, accumulator geometry (initCoords)                :501-541
 *   Appendix A linking: raster scan + Algorithm 5/6 greedy walk        :544-760
 *   cluster subdivision (Lowe's recursive segmentation)                :762-832
 *   Algorithm 2: kernel parameters, 2x2 eigen, heights, hmax           :834-1026  (eigen: base/math/compv_math_eigen.cxx:285-342)
 *   discard short kernels, Gmin, Gs                                    :1029-1062, :377
 *   Algorithm 4 voting (four quadrants, integer votes)                 :1065-1148
 *   3x3 smoothing + threshold, std::sort, sweep                        :1151-1247, :1282-1308
 * This file is C++ only because the peak ordering is std::sort's (libstdc++ introsort) tie order: the oracle must sort with the same
 * routine as the reference built by this toolchain.  Everything else is scalar arithmetic in the reference's operation order.
 *
 * Two details of the x86 build are reproduced on purpose (the compiled reference is what this oracle is pinned against):
 *  - kernel heights: for the first (n & -minpack) clusters the SSE2/AVX leaves evaluate 1/((sqrt(1-r^2)*s)*2pi), the scalar tail evaluates
 *    1/((2pi*s)*sqrt(1-r^2))  (intrin/x86/compv_core_feature_houghkht_intrin_avx.cxx:40-67 vs houghkht.cxx:858-884);
 *  - the peak scan: the SSE2 leaf covers rho_index 1 .. 4*ceil((rho_count-4)/4), the scalar tail is then called on &pcount[(rho_count&-4)+1] and
 *    reports its cells with indices RELATIVE to that pointer (houghkht.cxx:1176-1187).
 */
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

struct Pos { int y, x; double cy, cx; };
struct Range { size_t begin, end; };
struct Kernel { double rho, theta, h, sigma_theta_square, sigma_rho_square, m2, sigma_rho_times_theta; };
struct Vote { size_t rho_index, theta_index; int32_t count; };

struct Line { float rho, theta; size_t strength; }; // CompVHoughLine (compv_common.h:686-692)

inline double exp_small(double x) // houghkht.cxx:79-88: (1 + x/1024)^1024
{
    x = 1.0 + (x * (1.0 / 1024.0));
    for (int i = 0; i < 10; ++i) x *= x;
    return x;
}

// base/math/compv_math_eigen.cxx:285-342 (sort = true, norm = true)
void eigen2x2(const double A[4], double D[4], double Q[4])
{
    bool norm = true;
    const double trace = A[0] + A[3];
    const double half = trace / 2.0;
    const double det = (A[0] * A[3]) - (A[1] * A[2]);
    const double s = std::sqrt(((trace * trace) / 4.0) - det);
    D[1] = D[2] = 0.0;
    D[0] = half + s;
    D[3] = half - s;
    if (A[2] != 0) { Q[0] = D[0] - A[3]; Q[2] = A[2]; Q[1] = D[3] - A[3]; Q[3] = A[2]; }
    else if (A[1] != 0) { Q[0] = A[1]; Q[2] = D[0] - A[0]; Q[1] = A[1]; Q[3] = D[3] - A[0]; }
    else {
        norm = false;
        if (A[3] != 0.0) { Q[0] = 0.0; Q[2] = 1.0; Q[1] = 1.0; Q[3] = 0.0; }
        else { Q[0] = 1.0; Q[2] = 0.0; Q[1] = 0.0; Q[3] = 1.0; }
    }
    if (norm) {
        const double m02 = 1.0 / std::sqrt(Q[0] * Q[0] + Q[2] * Q[2]);
        const double m13 = 1.0 / std::sqrt(Q[1] * Q[1] + Q[3] * Q[3]);
        Q[0] *= m02; Q[2] *= m02; Q[1] *= m13; Q[3] *= m13;
    }
    if (D[0] < D[3]) {
        std::swap(Q[0], Q[1]); std::swap(Q[2], Q[3]); std::swap(D[0], D[3]);
    }
}

struct Kht {
    size_t W, H, stride;
    uint8_t* e; // private copy, erased while linking
    std::vector<Pos> poss;
    std::vector<Range> strings, clusters;
    double minDeviation; size_t minSize;

    // Algorithm 6: first set neighbour in the order TL, T, TR, L, R, BL, B, BR
    bool next(int& x, int& y) const
    {
        static const int dx[8] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 };
        for (int k = 0; k < 8; ++k) {
            const int nx = x + dx[k], ny = y + dy[k];
            if (nx < 0 || ny < 0 || nx >= (int)W || ny >= (int)H) continue;
            if (e[(size_t)ny * stride + nx]) { x = nx; y = ny; return true; }
        }
        return false;
    }

    void link(int xr, int yr) // Algorithm 5
    {
        const double hw = (double)W * 0.5, hh = (double)H * 0.5;
        const size_t begin = poss.size();
        int x = xr, y = yr;
        do { poss.push_back({ y, x, y - hh, x - hw }); e[(size_t)y * stride + x] = 0; } while (next(x, y));
        const size_t rev = poss.size();
        x = xr; y = yr;
        if (next(x, y)) {
            do { poss.push_back({ y, x, y - hh, x - hw }); e[(size_t)y * stride + x] = 0; } while (next(x, y));
        }
        const size_t end = poss.size();
        if (end - begin >= minSize) {
            std::reverse(poss.begin() + begin, poss.begin() + rev);
            strings.push_back({ begin, end });
        }
        else poss.resize(begin);
    }

    double subdivide(const Range& s, size_t a, size_t b) // houghkht.cxx:774-832
    {
        const size_t before = clusters.size();
        const Pos* p = &poss[s.begin];
        const int dxx = p[a].x - p[b].x, dyy = p[a].y - p[b].y;
        const double length = std::sqrt((double)((dxx * dxx) + (dyy * dyy)));
        size_t mi = a; int md = 0;
        for (size_t i = a + 1; i < b; ++i) {
            const int d = std::abs(((p[a].x - p[i].x) * dyy) - ((p[a].y - p[i].y) * dxx));
            if (d > md) { mi = i; md = d; }
        }
        const double ratio = length / std::max(((double)md / length), minDeviation);
        if ((mi - a + 1) >= minSize && (b - mi + 1) >= minSize) {
            const double rl = subdivide(s, a, mi);
            const double rr = subdivide(s, mi, b);
            if (rl > ratio || rr > ratio) return (rl > rr) ? rl : rr;
        }
        clusters.resize(before);
        clusters.push_back({ s.begin + a, s.begin + b + 1 });
        return ratio;
    }
};

} // namespace

// lines: array of {float rho, float theta, size_t strength}; returns the number of lines through *count (at most `capacity` are written).
// theta_deg_f is the value handed to CompVHough::newObj (degrees, float).  x86Simd != 0 reproduces the two x86 details described in the header.
ORC_API int orc_hough_kht(const uint8_t* edges, size_t w, size_t h, size_t stride, float rho_f, float theta_deg_f, size_t threshold,
                          float clusterMinDeviation, int clusterMinSize, float kernelMinHeight, int maxLines, int x86Simd,
                          void* lines_, size_t capacity, size_t* count, double* gsOut)
{
    Line* lines = static_cast<Line*>(lines_);
    *count = 0;
    if (gsOut) *gsOut = 1.0;
    if (!edges || !w || !h || stride < w || !(rho_f > 0.f) || rho_f > 1.f) return 20006;
    // ctor (houghkht.cxx:113-116)
    const float kPi = 3.1415926535897932384626433f;
    const float kPiOver180 = kPi / 180.f;
    const double dRho = (double)(rho_f * 1.f);
    const double dThetaRad = (double)(theta_deg_f * kPiOver180);
    // initCoords (houghkht.cxx:501-541)
    const double dThetaDeg = (dThetaRad * 180.0) / M_PI;
    const double r = std::sqrt((double)((w * w) + (h * h)));
    const size_t nRho = (size_t)((r + 1.0) / dRho);
    const size_t nTheta = (size_t)(180.0 / dThetaDeg);
    std::vector<double> rho(nRho + 1, 0.0), theta(nTheta + 1, 0.0);
    { double v = -(r * 0.5); for (size_t i = 1; i < nRho; ++i, v += dRho) rho[i] = v; }
    { double v = 0.0; for (size_t i = 1; i < nTheta; ++i, v += dThetaDeg) theta[i] = v; }
    const size_t cs = nRho + 2; // accumulator row pitch
    std::vector<int32_t> acc((nTheta + 2) * cs, 0);

    Kht k;
    k.W = w; k.H = h; k.stride = stride;
    k.minDeviation = (double)clusterMinDeviation; k.minSize = (size_t)clusterMinSize;
    std::vector<uint8_t> copy(edges, edges + stride * h);
    k.e = copy.data();
    // Appendix A: raster scan over the interior (houghkht.cxx:544-663)
    for (size_t y = 1; y + 1 < h; ++y) for (size_t x = 1; x + 1 < w; ++x) if (k.e[y * stride + x]) k.link((int)x, (int)y);
    if (k.strings.empty()) return 0;
    for (const Range& s : k.strings) k.subdivide(s, 0, (s.end - s.begin) - 1);
    if (k.clusters.empty()) return 0;

    // Algorithm 2 (houghkht.cxx:885-1026)
    const size_t n = k.clusters.size();
    std::vector<Kernel> kernels(n);
    std::vector<double> r0v(n), m0v(n), m2v(n), nsv(n);
    const double RAD2DEG = 180.0 / M_PI;
    for (size_t c = 0; c < n; ++c) {
        const Pos* pb = &k.poss[k.clusters[c].begin];
        const size_t np = k.clusters[c].end - k.clusters[c].begin;
        const double ns = 1.0 / (double)np;
        double mx = 0, my = 0;
        for (size_t i = 0; i < np; ++i) { mx += pb[i].cx; my += pb[i].cy; }
        mx *= ns; my *= ns;
        double cxx = 0, cyy = 0, cxy = 0;
        for (size_t i = 0; i < np; ++i) {
            const double cx = pb[i].cx - mx, cy = pb[i].cy - my;
            cxx += (cx * cx); cyy += (cy * cy); cxy += (cx * cy);
        }
        const double M[4] = { cxx, cxy, cxy, cyy };
        double D[4], Q[4];
        eigen2x2(M, D, Q);
        const double ux = Q[0], uy = Q[2];
        double vx = Q[1], vy = Q[3];
        if (vy < 0.0) { vx = -vx; vy = -vy; }
        kernels[c].rho = (vx * mx) + (vy * my);
        kernels[c].theta = std::acos(vx) * RAD2DEG;
        const double s1 = std::sqrt(1.0 - (vx * vx));
        m0v[c] = -(ux * mx) - (uy * my);
        m2v[c] = (s1 == 0.0) ? 0.0 : ((ux / s1) * RAD2DEG);
        double acc0 = 0.0;
        for (size_t i = 0; i < np; ++i) { const double t = (ux * (pb[i].cx - mx)) + (uy * (pb[i].cy - my)); acc0 += (t * t); }
        r0v[c] = acc0;
        nsv[c] = ns;
    }
    double hmax = 0.0;
    {
        const size_t pack = x86Simd ? (n >= 4 ? 4 : (n >= 2 ? 2 : 1)) : 1;
        const size_t simdCount = (pack > 1) ? (n & ~(pack - 1)) : 0;
        const double TWOPI = 2.0 * M_PI;
        for (size_t c = 0; c < n; ++c) {
            const double q0 = 1.0 / r0v[c];
            const double q1 = m0v[c] * q0, q2 = m2v[c] * q0;
            double srs = q1 * m0v[c] + nsv[c];
            const double srt = q1 * m2v[c];
            const double m2 = q2 * m0v[c];
            double sts = q2 * m2v[c];
            if (sts == 0.0) sts = 0.1;
            srs *= 4.0; sts *= 4.0;
            const double sst = std::sqrt(srs) * std::sqrt(sts);
            const double rr = srt / sst;
            const double omr = 1.0 - (rr * rr);
            const double height = (c < simdCount) ? (1.0 / ((std::sqrt(omr) * sst) * TWOPI)) : (1.0 / (TWOPI * sst * std::sqrt(omr)));
            kernels[c].sigma_rho_square = srs; kernels[c].sigma_rho_times_theta = srt; kernels[c].m2 = m2; kernels[c].sigma_theta_square = sts; kernels[c].h = height;
            hmax = std::max(hmax, height);
        }
    }
    // discard short kernels (houghkht.cxx:1029-1041)
    {
        const double scale = 1.0 / hmax, mh = (double)kernelMinHeight;
        kernels.erase(std::remove_if(kernels.begin(), kernels.end(), [=](const Kernel& q) { return (q.h * scale) < mh; }), kernels.end());
    }
    if (kernels.empty()) return 0;
    // Gmin (houghkht.cxx:1044-1062) with Eq15 (:834-847)
    double Gmin = DBL_MAX;
    for (const Kernel& q : kernels) {
        const double M[4] = { q.sigma_rho_square, q.sigma_rho_times_theta, q.m2, q.sigma_theta_square };
        double D[4], Q[4];
        eigen2x2(M, D, Q);
        const double r1 = std::sqrt(D[3]);
        const double rh = Q[1] * r1, th = Q[3] * r1;
        const double sst = std::sqrt(q.sigma_rho_square) * std::sqrt(q.sigma_theta_square);
        const double sc = 1.0 / sst;
        const double rr = q.sigma_rho_times_theta * sc;
        const double omr = 1.0 - (rr * rr);
        const double x = 1.0 / ((2.0 * M_PI) * sst * std::sqrt(omr));
        const double y = 1.0 / (2.0 * omr);
        const double z = ((rh * rh) / q.sigma_rho_square) - (((rr * 2.0) * rh * th) * sc) + ((th * th) / q.sigma_theta_square);
        const double g = x * exp_small(-z * y);
        if (g < Gmin) Gmin = g;
    }
    const double Gs = (Gmin == 0.0) ? 1.0 : std::max((1.0 / Gmin), 1.0);
    if (gsOut) *gsOut = Gs;

    // voting (houghkht.cxx:1065-1148)
    {
        const double rhoScale = 1.0 / dRho, thetaScale = 1.0 / dThetaDeg;
        const double rhoMaxNeg = rho[1];
        auto vote4 = [&](size_t rhoStartIdx, size_t thetaStartIdx, double rhoStart, double thetaStart, int incRhoIdx, int incThetaIdx, const Kernel& q) {
            const double incRho = dRho * incRhoIdx, incTheta = dThetaDeg * incThetaIdx;
            const double srsS = 1.0 / q.sigma_rho_square, stsS = 1.0 / q.sigma_theta_square;
            const double sst = std::sqrt(q.sigma_rho_square) * std::sqrt(q.sigma_theta_square);
            const double sc = 1.0 / sst;
            const double rr = q.sigma_rho_times_theta * sc;
            const double omr = 1.0 - (rr * rr);
            const double r2 = rr * 2.0;
            const double x = 1.0 / ((2.0 * M_PI) * sst * std::sqrt(omr));
            const double y = 1.0 / (2.0 * omr);
            size_t thetaIdx = thetaStartIdx, thetaCount = 0;
            double th = thetaStart, rh;
            do {
                if (!thetaIdx || thetaIdx > nTheta) {
                    rhoStartIdx = (nRho - rhoStartIdx) + 1;
                    thetaIdx = thetaIdx ? 1 : nTheta;
                    incRhoIdx = -incRhoIdx;
                }
                if (rhoStartIdx >= 1) {
                    int32_t* pc = &acc[thetaIdx * cs];
                    size_t rhoIdx = rhoStartIdx;
                    rh = rhoStart;
                    const double wv = (th * th) * stsS;
                    const double kk = r2 * th * sc;
                    double krho = kk * rh;
                    const double ki = kk * incRho;
                    double z = ((rh * rh) * srsS) - krho + wv;
                    int32_t votes;
                    while (rhoIdx <= nRho && (votes = (int32_t)(((x * exp_small(-z * y)) * Gs) + 0.5)) > 0) {
                        pc[rhoIdx] += votes;
                        rhoIdx += (size_t)(long long)incRhoIdx;
                        rh += incRho;
                        krho += ki;
                        z = ((rh * rh) * srsS) - krho + wv;
                    }
                    thetaIdx += (size_t)(long long)incThetaIdx;
                    th += incTheta;
                }
                else break;
            } while ((rh != rhoStart) && (++thetaCount < nTheta));
        };
        for (const Kernel& q : kernels) {
            const size_t ri = (size_t)(std::abs((q.rho - rhoMaxNeg) * rhoScale) + 0.5) + 1;
            const size_t ti = (size_t)(std::abs(q.theta * thetaScale) + 0.5) + 1;
            vote4(ri, ti, 0.0, 0.0, 1, 1, q);
            vote4(ri, ti - 1, 0.0, -dThetaDeg, 1, -1, q);
            vote4(ri - 1, ti, -dRho, 0.0, -1, 1, q);
            vote4(ri - 1, ti - 1, -dRho, -dThetaDeg, -1, -1, q);
        }
    }

    // peaks (houghkht.cxx:1151-1247, 1282-1308)
    std::vector<Vote> votes;
    {
        const int32_t thr = (int32_t)threshold;
        auto cell = [&](size_t ti, size_t ri, size_t reported) {
            const int32_t* c = &acc[ti * cs + ri];
            if (!(*c > 0)) return;
            const int32_t* t = c - cs; const int32_t* b = c + cs;
            const int32_t v = t[-1] + (t[0] << 1) + t[1] + b[-1] + (b[0] << 1) + b[1] + (c[-1] << 1) + (c[0] << 2) + (c[1] << 1);
            if (v >= thr) votes.push_back({ reported, ti, v });
        };
        for (size_t ti = 1; ti < nTheta; ++ti) {
            if (x86Simd && nRho > 4) {
                const size_t sseEnd = nRho - 3;
                size_t ri = 1;
                for (; ri < sseEnd; ri += 4) for (size_t j = 0; j < 4; ++j) cell(ti, ri + j, ri + j);
                const size_t consumed = (nRho & ~(size_t)3) + 1;
                const size_t remains = (nRho > consumed) ? (nRho - consumed) : 0;
                for (size_t rel = 1; rel < remains; ++rel) cell(ti, consumed + rel, rel);
            }
            else {
                for (size_t ri = 1; ri < nRho; ++ri) cell(ti, ri, ri);
            }
        }
    }
    std::sort(votes.begin(), votes.end(), [](const Vote& a, const Vote& b) { return a.count > b.count; });
    std::vector<uint8_t> visited((nTheta + 2) * cs, 0);
    size_t nLines = 0;
    const size_t lim = (maxLines <= 0) ? (size_t)INT_MAX : (size_t)maxLines;
    std::vector<Line> out;
    for (const Vote& v : votes) {
        uint8_t* pv = &visited[v.theta_index * cs + v.rho_index];
        const uint8_t* t = pv - cs; const uint8_t* b = pv + cs;
        const bool seen = t[-1] || t[0] || t[1] || pv[-1] || pv[1] || b[-1] || b[0] || b[1];
        if (!seen) out.push_back({ (float)rho[v.rho_index], (float)((theta[v.theta_index] * M_PI) / 180.0), (size_t)v.count });
        *pv = 0xff;
    }
    nLines = std::min(out.size(), lim);
    *count = nLines;
    for (size_t i = 0; i < nLines && i < capacity; ++i) lines[i] = out[i];
    return 0;
}
