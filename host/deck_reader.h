// Host-side mirror of the reference's input modules for the C++ driver (host/ns2dcomp.cpp):
// InputData::readInputData (dataLoader.f90:17-65) and MeshData::loadMeshData (dataLoader.f90:95-275).
// List-directed reads are reproduced with a whitespace/comma tokenizer that consumes whole lines per READ.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/cfdb.h"

namespace host {

struct Deck {
    std::string name;
    cfdb_params par{};
    int npoin = 0, nelem = 0;
    std::vector<double> X, Y;
    std::vector<int32_t> inpoel;
    std::vector<int32_t> ifixrho_node, ifixv_node, ifixt_node, wall, iset_n1, iset_n2, iset_elem, iset_id, ifm, i_m, master, slave;
    std::vector<double> rfixrho_value, rfixv_valuex, rfixv_valuey, rfixt_value;
    std::vector<unsigned char> smooth_fix;  // ns2DComp.ALE.f90:63-73
    int32_t print_flags[7] = {0, 0, 0, 0, 0, 0, 0};  // RHO VEL2 MACH PRES TEMP ENER POS
    cfdb_bc bc() const {
        cfdb_bc b{};
        b.nfixrho = (int)ifixrho_node.size(); b.ifixrho_node = ifixrho_node.data(); b.rfixrho_value = rfixrho_value.data();
        b.nfixv = (int)ifixv_node.size(); b.ifixv_node = ifixv_node.data(); b.rfixv_valuex = rfixv_valuex.data(); b.rfixv_valuey = rfixv_valuey.data();
        b.nwall = (int)wall.size() / 2; b.wall = wall.data();
        b.nfixt = (int)ifixt_node.size(); b.ifixt_node = ifixt_node.data(); b.rfixt_value = rfixt_value.data();
        b.nsets = (int)iset_id.size(); b.iset_n1 = iset_n1.data(); b.iset_n2 = iset_n2.data(); b.iset_elem = iset_elem.data(); b.iset_id = iset_id.data();
        b.nmove = (int)i_m.size(); b.i_m = i_m.data();
        b.nfix_move = (int)ifm.size(); b.ifm = ifm.data();
        return b;
    }
};

class Lines {
    std::ifstream f;
    std::string path;
public:
    explicit Lines(const std::string& p) : f(p), path(p) {
        if (!f) throw std::runtime_error("cannot open " + p);
    }
    void skip(int n = 1) { std::string s; for (int i = 0; i < n; ++i) std::getline(f, s); }
    std::vector<std::string> read() {  // one READ(1,*) statement = one line here
        std::string s;
        if (!std::getline(f, s)) throw std::runtime_error("unexpected end of " + path);
        for (auto& ch : s) if (ch == ',') ch = ' ';
        std::istringstream is(s);
        std::vector<std::string> t;
        std::string w;
        while (is >> w) t.push_back(w);
        return t;
    }
};
inline double num(std::string s) {
    for (auto& ch : s) if (ch == 'd' || ch == 'D') ch = 'e';
    return std::stod(s);
}

inline Deck read_deck(const std::string& dir) {
    Deck d;
    {
        std::ifstream e(dir + "/EULER.DAT");
        if (!e) throw std::runtime_error("cannot open " + dir + "/EULER.DAT");
        std::getline(e, d.name);
        while (!d.name.empty() && (d.name.back() == ' ' || d.name.back() == '\r')) d.name.pop_back();
    }
    cfdb_params& p = d.par;
    {
        Lines L(dir + "/" + d.name + "-1.dat");
        L.skip(); auto v = L.read();
        p.IRESTART = std::stoi(v.at(0)); p.MAXITER = std::stoi(v.at(1)); p.IPRINT = std::stoi(v.at(2)); p.MOVIE = std::stoi(v.at(3)); p.ITLOCAL = std::stoi(v.at(4));
        L.skip(); v = L.read();
        p.FSAFE = num(v.at(0)); p.U_inf = num(v.at(1)); p.V_inf = num(v.at(2)); p.MACH_inf = num(v.at(3)); p.T_inf = num(v.at(4)); p.RHO_inf = num(v.at(5)); p.P_inf = num(v.at(6));
        L.skip(); v = L.read();
        p.FMU = num(v.at(0)); p.FGX = num(v.at(1)); p.FGY = num(v.at(2)); p.QH = num(v.at(3));
        L.skip(); v = L.read();
        p.FK = num(v.at(0)); p.FR = num(v.at(1)); p.FCv = num(v.at(2)); p.GAMA = num(v.at(3)); p.NGAS = std::stoi(v.at(4));
        L.skip(); v = L.read();
        p.CTE = num(v.at(0));
        L.skip(2); v = L.read();
        p.MOVING = std::stoi(v.at(0)); p.XREF[0] = num(v.at(1)); p.YREF[0] = num(v.at(2));
        // RHOCHAR VEL2CHAR MACHCHAR PRESCHAR TEMPCHAR ENERCHAR POSCHAR (dataLoader.f90:46-48): '.si.' switches a block on
        L.skip(2); v = L.read();
        for (int k = 0; k < 7 && k < (int)v.size(); ++k) d.print_flags[k] = v[k].find(".si.") != std::string::npos;
    }
    // dataLoader.f90:58-64
    p.CTE = 1.0 / p.CTE;
    if (p.T_inf == 0.0) p.T_inf = p.P_inf / (p.FR * p.RHO_inf);
    if (p.P_inf == 0.0) p.P_inf = p.RHO_inf * p.FR * p.T_inf;
    if (p.RHO_inf == 0.0) p.RHO_inf = p.P_inf / (p.FR * p.T_inf);
    p.C_inf = std::sqrt(p.GAMA * p.FR * p.T_inf);
    if (std::sqrt(p.U_inf * p.U_inf + p.V_inf * p.V_inf) == 0.0) p.U_inf = p.C_inf * p.MACH_inf;

    // raw lists of <name>.dat: from the binary side-car <name>.cfdbmesh when present (same content, little-endian arrays:
    // cfd_b200/deck.py write_mesh_binary), else from the text file in dataLoader.f90's read order (:98-255)
    std::vector<int32_t> r_fixrho_n, r_fixvi_n, r_fixv_n, r_fixt_n;
    std::vector<double> r_fixrho_v, r_fixvi_x, r_fixvi_y, r_fixt_v;
    int nmaster = 0, nslave = 0;
    std::ifstream bin(dir + "/" + d.name + ".cfdbmesh", std::ios::binary);
    if (bin) {
        char magic[8];
        int32_t cnt[12];
        bin.read(magic, 8);
        bin.read(reinterpret_cast<char*>(cnt), sizeof cnt);
        if (!bin || std::string(magic, 8) != "CFDBMSH1") throw std::runtime_error("not a cfdb binary mesh: " + d.name + ".cfdbmesh");
        d.npoin = cnt[0]; d.nelem = cnt[1];
        nmaster = cnt[8]; nslave = cnt[9];
        auto rd_i = [&](std::vector<int32_t>& v, size_t n) { v.resize(n); bin.read(reinterpret_cast<char*>(v.data()), n * sizeof(int32_t)); };
        auto rd_d = [&](std::vector<double>& v, size_t n) { v.resize(n); bin.read(reinterpret_cast<char*>(v.data()), n * sizeof(double)); };
        rd_d(d.X, d.npoin); rd_d(d.Y, d.npoin); rd_i(d.inpoel, 3 * (size_t)d.nelem);
        rd_i(r_fixrho_n, cnt[2]); rd_d(r_fixrho_v, cnt[2]);
        rd_i(r_fixvi_n, cnt[3]); rd_d(r_fixvi_x, cnt[3]); rd_d(r_fixvi_y, cnt[3]);
        rd_i(r_fixv_n, cnt[4]);
        rd_i(d.wall, 2 * (size_t)cnt[5]);
        rd_i(r_fixt_n, cnt[6]); rd_d(r_fixt_v, cnt[6]);
        std::vector<int32_t> sets;
        rd_i(sets, 4 * (size_t)cnt[7]);
        for (int k = 0; k < cnt[7]; ++k) {
            d.iset_elem.push_back(sets[4 * k]); d.iset_n1.push_back(sets[4 * k + 1]); d.iset_n2.push_back(sets[4 * k + 2]); d.iset_id.push_back(sets[4 * k + 3]);
        }
        rd_i(d.master, cnt[8]); rd_i(d.slave, cnt[9]); rd_i(d.ifm, cnt[10]); rd_i(d.i_m, cnt[11]);
        if (!bin) throw std::runtime_error("truncated binary mesh: " + d.name + ".cfdbmesh");
    } else {
        Lines M(dir + "/" + d.name + ".dat");
        M.skip(); auto v = M.read();
        d.npoin = std::stoi(v.at(0)); d.nelem = std::stoi(v.at(1));
        M.skip(); v = M.read();
        int nfixrho = std::stoi(v.at(0)), nfixvi = std::stoi(v.at(1)), nfixv = std::stoi(v.at(2)), nwall = std::stoi(v.at(3)), nfixt = std::stoi(v.at(4)),
            nsets = std::stoi(v.at(5)), nfix_move = std::stoi(v.at(8)), nmove = std::stoi(v.at(9));
        nmaster = std::stoi(v.at(6)); nslave = std::stoi(v.at(7));
        M.skip(4);
        d.X.assign(d.npoin, 0.0); d.Y.assign(d.npoin, 0.0); d.inpoel.assign(3 * (size_t)d.nelem, 0);
        for (int k = 0; k < d.npoin; ++k) {
            v = M.read();
            int i = std::stoi(v.at(0));
            if (i < 1 || i > d.npoin) throw std::runtime_error("ERROR EN LA LECTURA DE NODOS");
            d.X[i - 1] = num(v.at(1)); d.Y[i - 1] = num(v.at(2));
        }
        M.skip();
        for (int k = 0; k < d.nelem; ++k) {
            v = M.read();
            int i = std::stoi(v.at(0));
            if (i < 1 || i > d.nelem) throw std::runtime_error("ERROR EN LA LECTURA DE ELEMENTOS");
            for (int j = 0; j < 3; ++j) d.inpoel[3 * (size_t)(i - 1) + j] = std::stoi(v.at(1 + j));
        }
        M.skip();
        for (int k = 0; k < nfixrho; ++k) { v = M.read(); r_fixrho_n.push_back(std::stoi(v.at(0))); r_fixrho_v.push_back(num(v.at(1))); }
        M.skip();
        for (int k = 0; k < nfixvi; ++k) { v = M.read(); r_fixvi_n.push_back(std::stoi(v.at(0))); r_fixvi_x.push_back(num(v.at(1))); r_fixvi_y.push_back(num(v.at(2))); }
        M.skip();
        for (int k = 0; k < nfixv; ++k) { v = M.read(); r_fixv_n.push_back(std::stoi(v.at(0))); }
        M.skip();
        for (int k = 0; k < nwall; ++k) { v = M.read(); d.wall.push_back(std::stoi(v.at(0))); d.wall.push_back(std::stoi(v.at(1))); }
        M.skip();
        for (int k = 0; k < nfixt; ++k) { v = M.read(); r_fixt_n.push_back(std::stoi(v.at(0))); r_fixt_v.push_back(num(v.at(1))); }
        M.skip();
        for (int k = 0; k < nsets; ++k) {
            v = M.read();
            d.iset_elem.push_back(std::stoi(v.at(0))); d.iset_n1.push_back(std::stoi(v.at(1))); d.iset_n2.push_back(std::stoi(v.at(2))); d.iset_id.push_back(std::stoi(v.at(3)));
        }
        if (nmaster != nslave) throw std::runtime_error("ERROR NODOS MASTER DISTINTO NODOS SLAVE");  // :223-226
        M.skip();
        for (int k = 0; k < nmaster; ++k) d.master.push_back(std::stoi(M.read().at(0)));
        M.skip();
        for (int k = 0; k < nmaster; ++k) d.slave.push_back(std::stoi(M.read().at(0)));
        M.skip();
        for (int k = 0; k < nfix_move; ++k) d.ifm.push_back(std::stoi(M.read().at(0)));
        M.skip();
        for (int k = 0; k < nmove; ++k) d.i_m.push_back(std::stoi(M.read().at(0)));
    }
    if (nmaster != nslave) throw std::runtime_error("ERROR NODOS MASTER DISTINTO NODOS SLAVE");  // :223-226
    // post-processing of the lists, dataLoader.f90:121-198
    for (size_t k = 0; k < r_fixrho_n.size(); ++k) {  // :121-130
        d.ifixrho_node.push_back(r_fixrho_n[k]);
        d.rfixrho_value.push_back(r_fixrho_v[k] < 0 ? 1.225 : r_fixrho_v[k] * p.RHO_inf);
    }
    for (size_t k = 0; k < r_fixvi_n.size(); ++k) {  // :136-143
        d.ifixv_node.push_back(r_fixvi_n[k]);
        d.rfixv_valuex.push_back(r_fixvi_x[k] * p.U_inf);
        d.rfixv_valuey.push_back(r_fixvi_y[k] * p.V_inf);
    }
    // TWALL is implicitly typed single precision (dataLoader.f90:147)
    float TWALL = (float)(p.T_inf * (1.0 + (p.GAMA - 1) / 2.0 * p.MACH_inf * p.MACH_inf));
    for (int n : r_fixv_n) {  // :152-161: no-slip nodes, then they also join the fixed-temperature list
        d.ifixv_node.push_back(n); d.rfixv_valuex.push_back(0.0); d.rfixv_valuey.push_back(0.0);
    }
    std::vector<int32_t> t_nodes;
    std::vector<double> t_vals;
    for (int n : r_fixv_n) { t_nodes.push_back(n); t_vals.push_back((double)TWALL); }
    for (size_t k = 0; k < r_fixt_n.size(); ++k) { t_nodes.push_back(r_fixt_n[k]); t_vals.push_back(r_fixt_v[k] * p.T_inf); }
    d.ifixt_node = t_nodes; d.rfixt_value = t_vals;
    d.smooth_fix.assign(d.npoin, 0);
    for (int n : d.i_m) d.smooth_fix.at(n - 1) = 1;
    for (int n : d.ifm) d.smooth_fix.at(n - 1) = 1;
    return d;
}

}  // namespace host
