// T, GAMM, RHO ... VX, VY, RMACH below are COLUMNS of the nodal record arrays (kernels.cuh: CF / WF), passed as the column's address.
// Relaxed-arithmetic translation unit: the same kernel templates as the exact build (kernels.cuh), compiled WITH FMA
// contraction (-fmad=true) under a different namespace.  Only the two kernels of the opt-in "fast" stage are
// instantiated here; everything else the library runs comes from cfdb.cu (-fmad=false).
#define CFDB_KNS kfast
#include "kernels.cuh"

namespace fastmode {

int launch_calcrhs_scatter(bool visc, bool ale, cudaStream_t st, int nelem, const int* inp, const double* U, const double* T,
                           const double* WX, const double* WY, const double* dNx, const double* dNy, const double* area,
                           const double* shoc, const double* dtl_arr, const double* dtl_sc, const double* ts1, const double* ts2,
                           const double* ts3, double Cv, double lambda_ref, double mu_ref, double gamma0, double T_inf, double cte,
                           double* RHS) {
    kfast::Gas g{Cv, lambda_ref, mu_ref, gamma0, T_inf, cte};
    const int B = 128, G = (nelem + B - 1) / B;
    if (visc && ale) kfast::calcrhs_scatter<true, true><<<G, B, 0, st>>>(nelem, inp, U, kfast::CF(T), WX, WY, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, RHS);
    else if (visc) kfast::calcrhs_scatter<true, false><<<G, B, 0, st>>>(nelem, inp, U, kfast::CF(T), WX, WY, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, RHS);
    else if (ale) kfast::calcrhs_scatter<false, true><<<G, B, 0, st>>>(nelem, inp, U, kfast::CF(T), WX, WY, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, RHS);
    else kfast::calcrhs_scatter<false, false><<<G, B, 0, st>>>(nelem, inp, U, kfast::CF(T), WX, WY, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, RHS);
    return (int)cudaGetLastError();
}

// the staged element kernel of the exact path, but compiled with FMA contraction (measurement of what -fmad=false costs)
int launch_calcrhs_staged_fma(bool visc, cudaStream_t st, int nelem, const int* inp, const double* U, const double* T,
                              const double* dNx, const double* dNy, const double* area, const double* shoc, const double* dtl_arr,
                              const double* dtl_sc, const double* ts1, const double* ts2, const double* ts3, double Cv,
                              double lambda_ref, double mu_ref, double gamma0, double T_inf, double cte, double* EC) {
    kfast::Gas g{Cv, lambda_ref, mu_ref, gamma0, T_inf, cte};
    const int B = 128, G = (nelem + B - 1) / B;
    if (visc) kfast::calcrhs_elem<true, false, false, 4><<<G, B, 0, st>>>(0, nelem, nelem, inp, U, nullptr, kfast::CF(T), nullptr, nullptr, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, EC, nullptr);
    else kfast::calcrhs_elem<false, false, false, 4><<<G, B, 0, st>>>(0, nelem, nelem, inp, U, nullptr, kfast::CF(T), nullptr, nullptr, dNx, dNy, area, shoc, dtl_arr, dtl_sc, ts1, ts2, ts3, g, EC, nullptr);
    return (int)cudaGetLastError();
}

int launch_node_update_rhs(cudaStream_t st, int npoin, const double* RHS, const double* U, const double* M, const double* GAMM,
                           const double* WX, const double* WY, const unsigned char* bcflag, int nb, const int* bnode,
                           const int* bkind, const double* bvx, const double* bvy, const double* brho, const double* bT,
                           const int* bwslot, const double* wnx, const double* wny, const int* wnvalid, double rk_fact, double FR,
                           double* U1, double* RHO, double* VX, double* VY, double* E, double* P, double* T, double* RMACH) {
    kfast::BcTab b;
    b.nb = nb; b.node = bnode; b.kind = bkind; b.vx = bvx; b.vy = bvy; b.rho = brho; b.Tfix = bT; b.wslot = bwslot;
    b.wn_x = wnx; b.wn_y = wny; b.wn_valid = wnvalid;
    kfast::node_update_rhs<<<(npoin + 255) / 256, 256, 0, st>>>(npoin, RHS, U, M, kfast::CF(GAMM), WX, WY, bcflag, b, rk_fact, FR, U1, kfast::WF(RHO),
                                                                  kfast::WF(VX), kfast::WF(VY), kfast::WF(E), kfast::WF(P), kfast::WF(T), kfast::WF(RMACH));
    return (int)cudaGetLastError();
}

}  // namespace fastmode
