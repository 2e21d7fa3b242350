// Init-time mesh optimiser of the reference (smoothing_mod::smoothing, smoothing.f90:21-370) for the C++ driver.
// Serial and order-dependent by construction (Gauss-Seidel sweeps with accept/reject), so it stays on the host:
// it runs once before the time loop (ns2DComp.ALE.f90:76) and changes X,Y for everything downstream.
// Same arithmetic order as the Fortran; bit-exact against the oracle (tests/test_host_driver.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../cfd_b200/csrc/host_topology.h"

namespace host {

class MeshSmoother {
    static constexpr double kTwoSqrt3 = 3.46410161513775;
    static constexpr int kMaxElemPerNode = 20, kNiter = 100, kMiter = 2, kNtry = 8;
    static constexpr double kTolMetric = .85, kFactorTolDist = 1e-2, kFactorDelta = 1e-2, kFactorPlus = 1.0, kFactorStep = 3.0;

    int npoin_, nelem_;
    const int32_t* inpoel_;
    double* X_;
    double* Y_;
    std::vector<int32_t> esup1_, esup2_, psup1_, psup2_;
    std::vector<char> smoothable_;
    double hmin_global_ = 1.0;
    // per-node work arrays of the optimiser (the reference keeps them as module variables: values of earlier nodes persist in the
    // entries a node with fewer elements does not overwrite, and getG differences all 20 entries)
    struct Work { double mu_vec[kMaxElemPerNode] = {0}, gx[kMaxElemPerNode] = {0}, gy[kMaxElemPerNode] = {0}; };
    Work serial_;

    void corners(int e1based, double* x, double* y) const {
        const int32_t* t = inpoel_ + 3 * (size_t)(e1based - 1);
        for (int i = 0; i < 3; ++i) { x[i] = X_[t[i] - 1]; y[i] = Y_[t[i] - 1]; }
    }
    static double twice_area(const double* x, const double* y) {
        return x[1] * y[2] + x[2] * y[0] + x[0] * y[1] - (x[1] * y[0] + x[2] * y[1] + x[0] * y[2]);
    }
    static double quality(const double* x, const double* y) {  // mu, smoothing.f90:305-317
        double a = twice_area(x, y);
        double l1 = (x[2] - x[1]) * (x[2] - x[1]) + (y[2] - y[1]) * (y[2] - y[1]);
        double l2 = (x[0] - x[2]) * (x[0] - x[2]) + (y[0] - y[2]) * (y[0] - y[2]);
        double l3 = (x[1] - x[0]) * (x[1] - x[0]) + (y[1] - y[0]) * (y[1] - y[0]);
        double l = l1 + l2 + l3;
        return kTwoSqrt3 * a / l;
    }
    static double size(const double* x, const double* y) {     // h, smoothing.f90:319-332
        double a = twice_area(x, y);
        double d1 = std::fabs(x[2] - x[1]) + std::fabs(y[2] - y[1]);
        double d2 = std::fabs(x[0] - x[2]) + std::fabs(y[0] - y[2]);
        double d3 = std::fabs(x[1] - x[0]) + std::fabs(y[1] - y[0]);
        double d = d1 + d2 + d3;
        return std::fabs(a) / d;
    }
    template <class F>
    void for_each_elem_of(int ip, F f) const {
        int base = esup2_[ip - 1];
        for (int k = base; k < esup2_[ip]; ++k) {
            double x[3], y[3];
            corners(esup1_[k], x, y);
            f(k - base, x, y);
        }
    }
    // getMu_vec, :173-199 (min_idx 1-based; strict '<' keeps the first minimum)
    void qualities(double* out, int ip, int* min_idx) const {
        if (min_idx) *min_idx = 1;
        double best = 1;
        for_each_elem_of(ip, [&](int j, const double* x, const double* y) {
            double m = quality(x, y);
            out[j] = m;
            if (min_idx && m < best) { best = m; *min_idx = j + 1; }
        });
    }
    double worst_quality(int ip) const {  // getMu_min, :201-219 ('<=')
        double r = 1;
        for_each_elem_of(ip, [&](int, const double* x, const double* y) { double m = quality(x, y); if (m <= r) r = m; });
        return r;
    }
    double local_size(int ip) const {     // getH_min, :274-291
        double r = 1;
        for_each_elem_of(ip, [&](int, const double* x, const double* y) { double h = size(x, y); if (h < r) r = h; });
        return r;
    }
    void gradient(int ip, Work& w) {      // getG, :221-243 — whole 20-entry arrays are differenced
        double* const mu_vec_ = w.mu_vec; double* const gx_ = w.gx; double* const gy_ = w.gy;
        double delta = hmin_global_ * kFactorDelta;
        double keep = X_[ip - 1];
        X_[ip - 1] = X_[ip - 1] + delta;
        qualities(gx_, ip, nullptr);
        for (int i = 0; i < kMaxElemPerNode; ++i) gx_[i] = (gx_[i] - mu_vec_[i]) / delta;
        X_[ip - 1] = keep;
        keep = Y_[ip - 1];
        Y_[ip - 1] = Y_[ip - 1] + delta;
        qualities(gy_, ip, nullptr);
        for (int i = 0; i < kMaxElemPerNode; ++i) gy_[i] = (gy_[i] - mu_vec_[i]) / delta;
        Y_[ip - 1] = keep;
    }
    double step_length(int min_idx, int ip, const Work& w) const {  // getStep, :245-272
        const double* const mu_vec_ = w.mu_vec; const double* const gx_ = w.gx; const double* const gy_ = w.gy;
        double gxm = gx_[min_idx - 1], gym = gy_[min_idx - 1], mum = mu_vec_[min_idx - 1];
        double g2 = gxm * gxm + gym * gym;
        double step = local_size(ip) * kFactorStep / (std::fabs(gxm) + std::fabs(gym));
        int n = esup2_[ip] - esup2_[ip - 1];
        for (int i = 0; i < n; ++i) {
            double gg = gxm * gx_[i] + gym * gy_[i];
            if (gg < 0) {
                double s1 = (mu_vec_[i] - mum) / (g2 - gg);
                if (s1 < step) step = s1;
            }
        }
        return step;
    }
    template <bool PAR>
    void move_node(int ip, int& min_idx, double& d_max, Work& w) {  // moveIpoin, :116-171
        double* const mu_vec_ = w.mu_vec; double* const gx_ = w.gx; double* const gy_ = w.gy;
        double x0 = X_[ip - 1], y0 = Y_[ip - 1];
        for (int it = 1; it <= kMiter; ++it) {
            double mu_min = mu_vec_[min_idx - 1];
            double xo = X_[ip - 1], yo = Y_[ip - 1];
            gradient(ip, w);
            double step = step_length(min_idx, ip, w);
            bool accepted = false;
            for (int j = 1; j <= kNtry; ++j) {
                double dx = step * gx_[min_idx - 1], dy = step * gy_[min_idx - 1];
                X_[ip - 1] = X_[ip - 1] + dx;   // not restored between tries, as written
                Y_[ip - 1] = Y_[ip - 1] + dy;
                int idx_new;
                qualities(mu_vec_, ip, &idx_new);
                if (mu_vec_[idx_new - 1] > mu_min * kFactorPlus) { min_idx = idx_new; accepted = true; break; }
                step = .5 * step;
            }
            if (!accepted) { X_[ip - 1] = xo; Y_[ip - 1] = yo; break; }
        }
        double ddx = X_[ip - 1] - x0, ddy = Y_[ip - 1] - y0;
        double d_move = ddx * ddx + ddy * ddy;
        if (d_move > d_max) d_max = d_move;
        if (d_move > std::numeric_limits<double>::min())  // update_list, :293-303
            for (int k = psup2_[ip - 1]; k < psup2_[ip]; ++k) {
                if (PAR) {   // nodes of one colour may flag a common neighbour at the same time (same value)
#pragma omp atomic write
                    smoothable_[psup1_[k] - 1] = 1;
                } else {
                    smoothable_[psup1_[k] - 1] = 1;
                }
            }
    }
    void laplacian_sweep(const unsigned char* fixed) {  // laplacianSmoothing, :74-114
        for (int ip = 1; ip <= npoin_; ++ip) {
            if (!smoothable_[ip - 1] || fixed[ip - 1]) continue;
            double xn = 0, yn = 0;
            double mu_old = worst_quality(ip);
            double xo = X_[ip - 1], yo = Y_[ip - 1];
            int n = psup2_[ip] - psup2_[ip - 1];
            for (int k = psup2_[ip - 1]; k < psup2_[ip]; ++k) { xn = xn + X_[psup1_[k] - 1]; yn = yn + Y_[psup1_[k] - 1]; }
            xn = xn / n; yn = yn / n;
            X_[ip - 1] = xn; Y_[ip - 1] = yn;
            if (worst_quality(ip) < mu_old) { X_[ip - 1] = xo; Y_[ip - 1] = yo; }
        }
    }

public:
    MeshSmoother(double* X, double* Y, const int32_t* inpoel, int npoin, int nelem)
        : npoin_(npoin), nelem_(nelem), inpoel_(inpoel), X_(X), Y_(Y), smoothable_(npoin, 0) {
        topo::build_esup(inpoel, nelem, npoin, esup1_, esup2_, nullptr);
        topo::build_psup(inpoel, npoin, esup1_, esup2_, psup1_, psup2_);
    }
    // smoothing, :21-72; returns the number of outer sweeps (0: nothing was smoothable)
    int run(const unsigned char* fixed) {
        bool any = false;
        hmin_global_ = 1.0;  // checkMesh, :351-370
        for (int e = 1; e <= nelem_; ++e) {
            double x[3], y[3];
            corners(e, x, y);
            if (quality(x, y) < kTolMetric) {
                for (int i = 0; i < 3; ++i) smoothable_[inpoel_[3 * (size_t)(e - 1) + i] - 1] = 1;
                any = true;
            }
            double h = size(x, y);
            if (h < hmin_global_) hmin_global_ = h;
        }
        if (!any) return 0;
        laplacian_sweep(fixed);
        laplacian_sweep(fixed);
        double tol_dist = kFactorTolDist * hmin_global_;
        int iter;
        for (iter = 1; iter <= kNiter; ++iter) {
            double d_max = 0.0;
            for (int ip = 1; ip <= npoin_; ++ip) {
                if (!smoothable_[ip - 1] || fixed[ip - 1]) continue;
                smoothable_[ip - 1] = 0;
                int min_idx;
                qualities(serial_.mu_vec, ip, &min_idx);
                if (serial_.mu_vec[min_idx - 1] < kTolMetric) move_node<false>(ip, min_idx, d_max, serial_);
            }
            if (d_max < tol_dist) break;  // squared distance against a length, as written (:62, :166)
        }
        return iter;
    }

    // a parallel variant (opt-in; NOT the reference's results).  The nodes are coloured greedily (ascending node
    // id, lowest colour not used by a node sharing an element) and every sweep visits the colours in order; the nodes of one
    // colour touch disjoint sets of elements whose other vertices do not move meanwhile, so they are optimised concurrently
    // (OpenMP) with the reference's own per-node procedure -- and the result does not depend on the number of threads.  The two
    // Laplacian pre-sweeps and the optimiser sweeps both run colour by colour; each node starts from zeroed work arrays.
    int run_colored(const unsigned char* fixed) {
        bool any = false;
        hmin_global_ = 1.0;
        for (int e = 1; e <= nelem_; ++e) {
            double x[3], y[3];
            corners(e, x, y);
            if (quality(x, y) < kTolMetric) {
                for (int i = 0; i < 3; ++i) smoothable_[inpoel_[3 * (size_t)(e - 1) + i] - 1] = 1;
                any = true;
            }
            double h = size(x, y);
            if (h < hmin_global_) hmin_global_ = h;
        }
        if (!any) return 0;
        // greedy node colouring over psup
        std::vector<int32_t> color((size_t)npoin_, -1);
        int ncol = 0;
        for (int ip = 1; ip <= npoin_; ++ip) {
            uint64_t used = 0;
            for (int k = psup2_[ip - 1]; k < psup2_[ip]; ++k) {
                int c = color[psup1_[k] - 1];
                if (c >= 0 && c < 64) used |= (uint64_t)1 << c;
            }
            int c = 0;
            while (c < 63 && (used >> c) & 1) ++c;
            color[ip - 1] = c;
            if (c + 1 > ncol) ncol = c + 1;
        }
        std::vector<int32_t> cptr((size_t)ncol + 1, 0), clist((size_t)npoin_);
        for (int ip = 0; ip < npoin_; ++ip) cptr[color[ip] + 1]++;
        for (int c = 0; c < ncol; ++c) cptr[c + 1] += cptr[c];
        {
            std::vector<int32_t> cur(cptr.begin(), cptr.end() - 1);
            for (int ip = 0; ip < npoin_; ++ip) clist[cur[color[ip]]++] = ip + 1;
        }
        auto laplacian_colored = [&]() {
            for (int c = 0; c < ncol; ++c) {
#pragma omp parallel for schedule(static)
                for (int q = cptr[c]; q < cptr[c + 1]; ++q) {
                    const int ip = clist[q];
                    if (!smoothable_[ip - 1] || fixed[ip - 1]) continue;
                    double xn = 0, yn = 0;
                    double mu_old = worst_quality(ip);
                    double xo = X_[ip - 1], yo = Y_[ip - 1];
                    int n = psup2_[ip] - psup2_[ip - 1];
                    for (int k = psup2_[ip - 1]; k < psup2_[ip]; ++k) { xn = xn + X_[psup1_[k] - 1]; yn = yn + Y_[psup1_[k] - 1]; }
                    xn = xn / n; yn = yn / n;
                    X_[ip - 1] = xn; Y_[ip - 1] = yn;
                    if (worst_quality(ip) < mu_old) { X_[ip - 1] = xo; Y_[ip - 1] = yo; }
                }
            }
        };
        laplacian_colored();
        laplacian_colored();
        double tol_dist = kFactorTolDist * hmin_global_;
        int iter;
        for (iter = 1; iter <= kNiter; ++iter) {
            double d_max = 0.0;
            for (int c = 0; c < ncol; ++c) {
                // the flags of this colour's nodes are read and cleared here; the flags the moves set belong to other colours
#pragma omp parallel for schedule(dynamic, 256) reduction(max : d_max)
                for (int q = cptr[c]; q < cptr[c + 1]; ++q) {
                    const int ip = clist[q];
                    if (!smoothable_[ip - 1] || fixed[ip - 1]) continue;
                    smoothable_[ip - 1] = 0;
                    Work w;
                    int min_idx;
                    qualities(w.mu_vec, ip, &min_idx);
                    if (w.mu_vec[min_idx - 1] < kTolMetric) move_node<true>(ip, min_idx, d_max, w);
                }
            }
            if (d_max < tol_dist) break;
        }
        return iter;
    }
};

}  // namespace host
