// ns2dcomp — C++ host driver mirroring PROGRAM NSComp2D (ns2DComp.ALE.f90:8-389) on top of the C ABI of
// libcfdb200.so.  Written in C++ because the reference's toolchain (Fortran) does not exist in the build
// image; call for call it is the reference's program:
//   readInputData / loadMeshData      -> host::read_deck                     (dataLoader.f90)
//   SMOOTH_FIX + smoothing            -> host::MeshSmoother                  (ns2DComp.ALE.f90:63-76)
//   RESTART, NORMALES, DERIV, MASAS, laplace -> cfdb_create + cfdb_init      (:59-100)
//   time loop                         -> cfdb_step, one pass per iteration   (:138-282)
//   print steps                       -> residual norms, <name>.cnv, run-info block, FUSIBLE abort (:186-224), GiD post file,
//                                        SKIN.DAT, DESPLAZAMIENTO, FORCES, <name>.RST (:225-254); IRESTART = 1 reads <name>.RST (:423-431)
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

// <name>.RST, Fortran unformatted sequential (PRINTREST ns2DComp.ALE.f90:898-917, RESTART :423-431): record 1 = (ITER int32,
// TIME real64), then one record per node = (U(1:4), T, GAMM); 4-byte length markers around every record
static bool write_rst(const std::string& path, int iter, double time, const std::vector<double>& U, const std::vector<double>& T,
                      const std::vector<double>& G) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    int32_t l = 12;
    std::fwrite(&l, 4, 1, f); std::fwrite(&iter, 4, 1, f); std::fwrite(&time, 8, 1, f); std::fwrite(&l, 4, 1, f);
    l = 48;
    for (size_t n = 0; n < T.size(); ++n) {
        std::fwrite(&l, 4, 1, f); std::fwrite(&U[4 * n], 8, 4, f); std::fwrite(&T[n], 8, 1, f); std::fwrite(&G[n], 8, 1, f); std::fwrite(&l, 4, 1, f);
    }
    std::fclose(f);
    return true;
}
static bool read_rst(const std::string& path, size_t npoin, std::vector<double>& U, std::vector<double>& T, std::vector<double>& G) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    int32_t l0 = 0, l1 = 0, it = 0;
    double time = 0;
    bool ok = std::fread(&l0, 4, 1, f) == 1 && std::fread(&it, 4, 1, f) == 1 && std::fread(&time, 8, 1, f) == 1 && std::fread(&l1, 4, 1, f) == 1 &&
              l0 == 12 && l1 == 12;
    U.resize(4 * npoin); T.resize(npoin); G.resize(npoin);
    for (size_t n = 0; ok && n < npoin; ++n)
        ok = std::fread(&l0, 4, 1, f) == 1 && l0 == 48 && std::fread(&U[4 * n], 8, 4, f) == 4 && std::fread(&T[n], 8, 1, f) == 1 &&
             std::fread(&G[n], 8, 1, f) == 1 && std::fread(&l1, 4, 1, f) == 1 && l1 == 48;
    std::fclose(f);
    return ok;
}

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
    if (d.par.IRESTART == 1) {
        // RESTART with IRESTART == 1 (ns2DComp.ALE.f90:423-431): U, T, GAMM from <name>.RST; ITER / TIME are read and dropped,
        // VEL_X / VEL_Y are not restored (left at zero, what fresh ALLOCATE memory holds), exactly as the reference does
        std::vector<double> U, T, G, z((size_t)d.npoin, 0.0);
        if (!read_rst(dir + "/" + d.name + ".RST", (size_t)d.npoin, U, T, G)) {
            std::fprintf(stderr, "IRESTART = 1 but %s/%s.RST is missing or does not hold %d node records\n", dir.c_str(), d.name.c_str(), d.npoin);
            return 1;
        }
        if (cfdb_set(ctx, "U", U.data(), (int64_t)U.size()) || cfdb_set(ctx, "T", T.data(), d.npoin) || cfdb_set(ctx, "GAMM", G.data(), d.npoin) ||
            cfdb_set(ctx, "VEL_X", z.data(), d.npoin) || cfdb_set(ctx, "VEL_Y", z.data(), d.npoin))
            die("restart");
        std::printf(" RESTART: state read from %s.RST\n", d.name.c_str());
    }
    std::printf("****-------> RUNGE-KUTTA DE  4  ORDEN <-------****\n\n");
    std::FILE* cnv = std::fopen((dir + "/" + d.name + ".cnv").c_str(), "w");
    int iter = 0, iterprint = 0;
    bool flavia_written = false, desp_written = false;
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
            // the other files of a print step (:228-254): SKIN.DAT (FORCE_VISC ran inside cfdb_step when FMU /= 0), one record of
            // DESPLAZAMIENTO, FORCES, and the restart dump (U and T of the same step: the reference pairs the step-start U with the
            // step-end T because it calls PRINTREST before its U = U1)
            if (d.par.FMU != 0.0 && cfdb_write_skin(ctx, (dir + "/SKIN.DAT").c_str())) die("cfdb_write_skin");
            if (cfdb_write_desplazamiento(ctx, (dir + "/DESPLAZAMIENTO").c_str(), time, desp_written)) die("cfdb_write_desplazamiento");
            desp_written = true;
            if (cfdb_write_forces(ctx, (dir + "/FORCES").c_str())) die("cfdb_write_forces");
            {
                std::vector<double> U(4 * (size_t)d.npoin), T((size_t)d.npoin), G((size_t)d.npoin);
                if (cfdb_get(ctx, "U", U.data(), (int64_t)U.size()) || cfdb_get(ctx, "T", T.data(), d.npoin) || cfdb_get(ctx, "GAMM", G.data(), d.npoin))
                    die("cfdb_get (PRINTREST)");
                if (!write_rst(dir + "/" + d.name + ".RST", iter, time, U, T, G)) { std::fprintf(stderr, "cannot write %s.RST\n", d.name.c_str()); return 1; }
            }
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
