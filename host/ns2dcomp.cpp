// ns2dcomp — C++ host driver mirroring PROGRAM NSComp2D (ns2DComp.ALE.f90:8-389) on top of the C ABI of
// libcfdb200.so.  Written in C++ because the reference's toolchain (Fortran) does not exist in the build
// image; call for call it is the reference's program:
//   readInputData / loadMeshData      -> host::read_deck                     (dataLoader.f90)
//   SMOOTH_FIX + smoothing            -> host::MeshSmoother                  (ns2DComp.ALE.f90:63-76)
//   RESTART, NORMALES, DERIV, MASAS, laplace -> cfdb_create + cfdb_init      (:59-100)
//   time loop                         -> cfdb_step, one pass per iteration   (:138-282)
//   print steps                       -> residual norms, <name>.cnv, run-info block, FUSIBLE abort (:186-224)
// Usage:  ns2dcomp [case_dir] [--check-deck] [--no-smoothing] [--device N] [--dump FIELD:FILE ...]
// The .cnv line is ITER TIME r1 r2 r3 r4 on ONE line ('(I7,5E14.6)'): the reference's format has one slot too
// few.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/cfdb.h"
#include "deck_reader.h"
#include "mesh_smoothing.h"
#include "fortran_format.h"

static void die(const char* who) {
    std::fprintf(stderr, "%s: %s\n", who, cfdb_last_error());
    std::exit(1);  // the reference's error convention is STOP
}

int main(int argc, char** argv) {
    std::string dir = ".";
    bool check_only = false, smooth = true;
    int device = 0;
    std::vector<std::pair<std::string, std::string>> dumps;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--check-deck") check_only = true;
        else if (a == "--no-smoothing") smooth = false;
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--dump" && i + 1 < argc) {
            std::string s = argv[++i];
            auto c = s.find(':');
            dumps.push_back({s.substr(0, c), s.substr(c + 1)});
        } else dir = a;
    }
    host::Deck d;
    try {
        d = host::read_deck(dir);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::printf(" TOTAL NODOS LEIDOS:%d\n TOTAL ELEMENTOS LEIDOS:%d\n", d.npoin, d.nelem);
    std::printf(" nfixrho=%zu nfixv=%zu nwall=%zu nfixt=%zu nsets=%zu nfix_move=%zu nmove=%zu\n", d.ifixrho_node.size(),
                d.ifixv_node.size(), d.wall.size() / 2, d.ifixt_node.size(), d.iset_id.size(), d.ifm.size(), d.i_m.size());
    int sweeps = 0;
    if (smooth) {
        host::MeshSmoother sm(d.X.data(), d.Y.data(), d.inpoel.data(), d.npoin, d.nelem);
        sweeps = sm.run(d.smooth_fix.data());
        std::printf(" =============SMOOTHING============= sweeps:%d\n", sweeps);
    }
    if (check_only) {
        for (auto& dm : dumps) {  // X / Y after smoothing, raw float64
            const std::vector<double>& v = dm.first == "X" ? d.X : d.Y;
            std::FILE* f = std::fopen(dm.second.c_str(), "wb");
            std::fwrite(v.data(), sizeof(double), v.size(), f);
            std::fclose(f);
        }
        double sx = 0, sy = 0;
        for (int i = 0; i < d.npoin; ++i) { sx += d.X[i]; sy += d.Y[i]; }
        std::printf(" deck ok: U_inf=%.17g CTE=%.17g sumX=%.17g sumY=%.17g\n", d.par.U_inf, d.par.CTE, sx, sy);
        return 0;
    }
    auto t0 = std::chrono::steady_clock::now();
    cfdb_ctx* ctx = nullptr;
    cfdb_bc bc = d.bc();
    if (cfdb_create(&ctx, &d.par, d.npoin, d.nelem, d.X.data(), d.Y.data(), d.inpoel.data(), &bc, device)) die("cfdb_create");
    if (cfdb_init(ctx)) die("cfdb_init");
    std::printf("****-------> RUNGE-KUTTA DE  4  ORDEN <-------****\n\n");
    std::FILE* cnv = std::fopen((dir + "/" + d.name + ".cnv").c_str(), "w");
    int iter = 0, iterprint = 0;
    bool flavia_written = false;
    const int MAXITER = d.par.MAXITER, IPRINT = d.par.IPRINT;
    while (iter < MAXITER) {  // ns2DComp.ALE.f90:138
        iter += 1;
        if (cfdb_step(ctx, 1)) die("cfdb_step");
        iterprint += 1;
        if (iterprint == IPRINT || iter == MAXITER) {  // :186
            double er[4], err[4], time, dtmin;
            if (cfdb_step_norms(ctx, er, err)) die("cfdb_step_norms");
            cfdb_get_scalar(ctx, "TIME", &time);
            cfdb_get_scalar(ctx, "DTMIN", &dtmin);
            double r[4], fus = 0;
            for (int i = 0; i < 4; ++i) { r[i] = std::sqrt(er[i] / err[i]); if (r[i] > fus) fus = r[i]; }
            std::fprintf(cnv, "%s\n", ffmt::cnv_record(iter, time, r).c_str());
            std::fflush(cnv);
            if (fus > 1.e2) {  // FUSIBLE, :202-210
                std::printf("      ERROR CONVERGENCIA\n    *****  OVERFLOW  *****\n");
                return 2;
            }
            std::printf("CCCC  ----> INFORMACION DE LA CORRIDA <----  CCCC\nPASOS EJECUTADOS:%6d\nTIEMPO ACUMULADO:%12.4E\nPASO DE TIEMPO:%12.4E\n",
                        iter, time, dtmin);
            std::printf("Continuidad %12.4E\nMomento u   %12.4E\nMomento v   %12.4E\nEnergia     %12.4E\n\n", r[0], r[1], r[2], r[3]);
            // PRINTFLAVIA (:225-226): with MOVIE=0 the file is reopened and rewritten at every print step (:709-711)
            if (cfdb_printflavia(ctx, (dir + "/" + d.name + ".flavia.res").c_str(), iter, d.print_flags,
                                 d.par.MOVIE == 1 && flavia_written))
                die("cfdb_printflavia");
            flavia_written = true;
            iterprint = 0;
        }
    }
    if (cfdb_sync(ctx)) die("cfdb_sync");
    for (auto& dm : dumps) {  // raw float64 dumps for tests / post-processing
        long n = cfdb_field_size(ctx, dm.first.c_str());
        if (n < 0) { std::fprintf(stderr, "unknown field %s\n", dm.first.c_str()); return 1; }
        std::vector<double> buf(n);
        if (cfdb_get(ctx, dm.first.c_str(), buf.data(), n)) die("cfdb_get");
        std::FILE* f = std::fopen(dm.second.c_str(), "wb");
        std::fwrite(buf.data(), sizeof(double), n, f);
        std::fclose(f);
    }
    std::fclose(cnv);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(" TIEMPO TOTAL: %g\n\n****-------> FIN DEL CALCULO <-------****\n", secs);
    cfdb_destroy(ctx);
    return 0;
}
