// Host-side gate table and lowering of (name | matrix, wires) into what the kernels consume.
//
// Gate DEFINITIONS (the parity contract) follow the reference's
//   simulator/cuGates_host.hpp:27-1326      (matrices, exp(-i theta/2 P) convention)
//   simulator/StateVectorCudaManaged.hpp:321-560  (leading wires = controls, last wire(s) = target)
//   simulator/StateVectorCudaBase.hpp:288-312     (gate name -> number of controls)
//   algorithms/GateGenerators.hpp:58-321, algorithms/AdjointDiffGPU.hpp:96-114 (generators, scaling)
// The REPRESENTATION is our own: instead of dense matrices handed to a library, every gate is lowered
// to the cheapest of {diagonal phase table, parity phase, two-level rotation, dense 2^k block} with
// controls folded into an index predicate, so that a gate only touches the amplitudes it changes.
#include <algorithm>
#include <cmath>
#include <map>

#include "qsv_internal.h"

namespace qsv {

namespace {

const cplx I1(0.0, 1.0);

using Mat = std::vector<cplx>;

Mat mat2(cplx a, cplx b, cplx c, cplx d) { return {a, b, c, d}; }

Mat m_rx(double t) {
    double c = std::cos(t / 2), s = std::sin(t / 2);
    return mat2(c, -I1 * s, -I1 * s, c);
}
Mat m_ry(double t) {
    double c = std::cos(t / 2), s = std::sin(t / 2);
    return mat2(c, -s, s, c);
}
Mat m_rot(double phi, double theta, double omega) {
    // RZ(omega) RY(theta) RZ(phi)
    double c = std::cos(theta / 2), s = std::sin(theta / 2);
    cplx ep = std::polar(1.0, (phi + omega) / 2), em = std::polar(1.0, (phi - omega) / 2);
    return mat2(std::conj(ep) * c, -em * s, std::conj(em) * s, ep * c);
}

Mat dagger(const Mat &m, int dim) {
    Mat r(m.size());
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) r[(size_t)i * dim + j] = std::conj(m[(size_t)j * dim + i]);
    return r;
}

Mat kron(const Mat &a, int da, const Mat &b, int db) {
    int d = da * db;
    Mat r((size_t)d * d);
    for (int i = 0; i < da; ++i)
        for (int j = 0; j < da; ++j)
            for (int k = 0; k < db; ++k)
                for (int l = 0; l < db; ++l)
                    r[(size_t)(i * db + k) * d + (j * db + l)] = a[(size_t)i * da + j] * b[(size_t)k * db + l];
    return r;
}

const Mat PX = mat2(0, 1, 1, 0);
const Mat PY = mat2(0, -I1, I1, 0);
const Mat PZ = mat2(1, 0, 0, -1);

// exp(-i t/2 P (x) P)
Mat m_ising(const Mat &p, double t) {
    Mat pp = kron(p, 2, p, 2);
    double c = std::cos(t / 2), s = std::sin(t / 2);
    Mat r(16);
    for (int i = 0; i < 16; ++i) r[i] = -I1 * s * pp[i];
    for (int i = 0; i < 4; ++i) r[i * 4 + i] += c;
    return r;
}

enum Core {
    C_ID, C_X, C_Y, C_Z, C_H, C_S, C_T, C_RX, C_RY, C_RZ, C_PHASE, C_ROT,
    C_SWAP, C_IXX, C_IYY, C_IZZ, C_SE, C_SEM, C_SEP, C_DE, C_DEM, C_DEP, C_MRZ
};

struct Entry {
    GateInfo info;
    int n_ctrl;  // leading wires that are controls
    Core core;
};

const Entry TABLE[] = {
    {{"Identity", 1, 0}, 0, C_ID},
    {{"PauliX", 1, 0}, 0, C_X},
    {{"PauliY", 1, 0}, 0, C_Y},
    {{"PauliZ", 1, 0}, 0, C_Z},
    {{"Hadamard", 1, 0}, 0, C_H},
    {{"S", 1, 0}, 0, C_S},
    {{"T", 1, 0}, 0, C_T},
    {{"RX", 1, 1}, 0, C_RX},
    {{"RY", 1, 1}, 0, C_RY},
    {{"RZ", 1, 1}, 0, C_RZ},
    {{"PhaseShift", 1, 1}, 0, C_PHASE},
    {{"Rot", 1, 3}, 0, C_ROT},
    {{"CNOT", 2, 0}, 1, C_X},
    {{"CY", 2, 0}, 1, C_Y},
    {{"CZ", 2, 0}, 1, C_Z},
    {{"SWAP", 2, 0}, 0, C_SWAP},
    {{"IsingXX", 2, 1}, 0, C_IXX},
    {{"IsingYY", 2, 1}, 0, C_IYY},
    {{"IsingZZ", 2, 1}, 0, C_IZZ},
    {{"CRX", 2, 1}, 1, C_RX},
    {{"CRY", 2, 1}, 1, C_RY},
    {{"CRZ", 2, 1}, 1, C_RZ},
    {{"CRot", 2, 3}, 1, C_ROT},
    {{"ControlledPhaseShift", 2, 1}, 1, C_PHASE},
    {{"SingleExcitation", 2, 1}, 0, C_SE},
    {{"SingleExcitationMinus", 2, 1}, 0, C_SEM},
    {{"SingleExcitationPlus", 2, 1}, 0, C_SEP},
    {{"Toffoli", 3, 0}, 2, C_X},
    {{"CSWAP", 3, 0}, 1, C_SWAP},
    {{"DoubleExcitation", 4, 1}, 0, C_DE},
    {{"DoubleExcitationMinus", 4, 1}, 0, C_DEM},
    {{"DoubleExcitationPlus", 4, 1}, 0, C_DEP},
    {{"MultiRZ", 0, 1}, 0, C_MRZ},
};

const Entry *find_entry(const std::string &name) {
    for (const auto &e : TABLE)
        if (name == e.info.name) return &e;
    return nullptr;
}

int bit_of(int n, int wire) {
    QSV_CHECK(wire >= 0 && wire < n, "wire index " + std::to_string(wire) + " out of range for " +
                                         std::to_string(n) + " qubits");
    return n - 1 - wire;
}

void check_distinct(const std::vector<int> &wires) {
    for (size_t i = 0; i < wires.size(); ++i)
        for (size_t j = i + 1; j < wires.size(); ++j)
            QSV_CHECK(wires[i] != wires[j], "repeated wire in gate");
}

// dense block on tgt_bits (MSB first) with control mask
LoweredGate make_dense(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, Mat m) {
    LoweredGate g;
    g.kind = LoweredGate::DENSE;
    g.k = (int)tgt_bits.size();
    g.tgt_bits = tgt_bits;
    g.ctrl_mask = ctrl_mask;
    g.holes = tgt_bits;
    for (int b = 0; b < 64; ++b)
        if (ctrl_mask >> b & 1) g.holes.push_back(b);
    std::sort(g.holes.begin(), g.holes.end());
    int dim = 1 << g.k;
    g.offs.resize(dim);
    for (int j = 0; j < dim; ++j) {
        uint64_t o = 0;
        for (int b = 0; b < g.k; ++b)
            if (j >> (g.k - 1 - b) & 1) o |= 1ull << tgt_bits[b];
        g.offs[j] = o;
    }
    g.mat = std::move(m);
    return g;
}

// 2x2 block acting on the two basis states p, q (indices in wire order over tgt_bits) of a
// k-target gate that is the identity elsewhere: touches only 2 / 2^k of the amplitudes.
LoweredGate make_two_level(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, int p, int q, Mat m2) {
    LoweredGate full = make_dense(tgt_bits, ctrl_mask, {});
    LoweredGate g;
    g.kind = LoweredGate::DENSE;
    g.k = 1;
    g.ctrl_mask = ctrl_mask;
    g.holes = full.holes;
    g.offs = {full.offs[p], full.offs[q]};
    g.tgt_bits.clear();  // no single target bit: the pair is (base+offs[0], base+offs[1])
    g.mat = std::move(m2);
    return g;
}

LoweredGate make_phase(uint64_t ctrl_mask, cplx ph) {
    LoweredGate g;
    g.kind = LoweredGate::DIAG;
    g.k = 0;
    g.ctrl_mask = ctrl_mask;
    g.mat = {ph};
    return g;
}

LoweredGate make_diag(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, Mat d) {
    LoweredGate g;
    g.kind = LoweredGate::DIAG;
    g.k = (int)tgt_bits.size();
    g.tgt_bits = tgt_bits;
    g.ctrl_mask = ctrl_mask;
    g.mat = std::move(d);
    return g;
}

LoweredGate make_parity(uint64_t zmask, uint64_t ctrl_mask, cplx even, cplx odd) {
    LoweredGate g;
    g.kind = LoweredGate::PARITY;
    g.zmask = zmask;
    g.ctrl_mask = ctrl_mask;
    g.mat = {even, odd};
    return g;
}

Mat m_single_exc(double t, int phase) {
    double c = std::cos(t / 2), s = std::sin(t / 2);
    cplx e = std::polar(1.0, phase * t / 2);
    Mat m(16, 0.0);
    m[0] = m[15] = e;
    m[5] = m[10] = c;
    m[6] = -s;
    m[9] = s;
    return m;
}

Mat m_double_exc(double t, int phase) {
    double c = std::cos(t / 2), s = std::sin(t / 2);
    cplx e = std::polar(1.0, phase * t / 2);
    Mat m(256, 0.0);
    for (int i = 0; i < 16; ++i) m[i * 16 + i] = e;
    m[3 * 16 + 3] = m[12 * 16 + 12] = c;
    m[3 * 16 + 12] = -s;
    m[12 * 16 + 3] = s;
    return m;
}

Mat core_matrix(Core core, const std::vector<double> &p, int n_tgt) {
    const double r = 1.0 / std::sqrt(2.0);
    switch (core) {
    case C_ID: return mat2(1, 0, 0, 1);
    case C_X: return PX;
    case C_Y: return PY;
    case C_Z: return PZ;
    case C_H: return mat2(r, r, r, -r);
    case C_S: return mat2(1, 0, 0, I1);
    case C_T: return mat2(1, 0, 0, std::polar(1.0, M_PI / 4));
    case C_RX: return m_rx(p[0]);
    case C_RY: return m_ry(p[0]);
    case C_RZ: return mat2(std::polar(1.0, -p[0] / 2), 0, 0, std::polar(1.0, p[0] / 2));
    case C_PHASE: return mat2(1, 0, 0, std::polar(1.0, p[0]));
    case C_ROT: return m_rot(p[0], p[1], p[2]);
    case C_SWAP: {
        Mat m(16, 0.0);
        m[0] = m[6] = m[9] = m[15] = 1;
        return m;
    }
    case C_IXX: return m_ising(PX, p[0]);
    case C_IYY: return m_ising(PY, p[0]);
    case C_IZZ: return m_ising(PZ, p[0]);
    case C_SE: return m_single_exc(p[0], 0);
    case C_SEM: return m_single_exc(p[0], -1);
    case C_SEP: return m_single_exc(p[0], +1);
    case C_DE: return m_double_exc(p[0], 0);
    case C_DEM: return m_double_exc(p[0], -1);
    case C_DEP: return m_double_exc(p[0], +1);
    case C_MRZ: {
        int dim = 1 << n_tgt;
        Mat m((size_t)dim * dim, 0.0);
        for (int i = 0; i < dim; ++i)
            m[(size_t)i * dim + i] = std::polar(1.0, (__builtin_popcount(i) & 1) ? p[0] / 2 : -p[0] / 2);
        return m;
    }
    }
    fail("internal: unknown gate core");
}

Mat controlled(const Mat &u, int du, int n_ctrl) {
    int dim = du << n_ctrl;
    Mat m((size_t)dim * dim, 0.0);
    for (int i = 0; i < dim - du; ++i) m[(size_t)i * dim + i] = 1;
    for (int i = 0; i < du; ++i)
        for (int j = 0; j < du; ++j) m[(size_t)(dim - du + i) * dim + (dim - du + j)] = u[(size_t)i * du + j];
    return m;
}

}  // namespace

LoweredGate make_dense_gate(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, std::vector<cplx> m) {
    return make_dense(tgt_bits, ctrl_mask, std::move(m));
}
LoweredGate make_diag_gate(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, std::vector<cplx> d) {
    return make_diag(tgt_bits, ctrl_mask, std::move(d));
}

const GateInfo *find_gate(const std::string &name) {
    const Entry *e = find_entry(name);
    return e ? &e->info : nullptr;
}

std::vector<cplx> named_gate_matrix(const std::string &name, const std::vector<double> &params,
                                    int n_wires) {
    const Entry *e = find_entry(name);
    QSV_CHECK(e != nullptr, "Currently unsupported gate: " + name);
    QSV_CHECK((int)params.size() >= e->info.n_params, "gate " + name + " needs " +
                                                          std::to_string(e->info.n_params) + " parameter(s)");
    int nw = e->info.n_wires ? e->info.n_wires : n_wires;
    int n_tgt = nw - e->n_ctrl;
    Mat u = core_matrix(e->core, params, n_tgt);
    return controlled(u, 1 << n_tgt, e->n_ctrl);
}

LoweredGate lower_named(int n, const std::string &name, const std::vector<int> &wires,
                        const std::vector<double> &params, bool adjoint) {
    const Entry *e = find_entry(name);
    QSV_CHECK(e != nullptr, "Currently unsupported gate: " + name);
    if (e->info.n_wires != 0)
        QSV_CHECK((int)wires.size() == e->info.n_wires,
                  "gate " + name + " acts on " + std::to_string(e->info.n_wires) + " wire(s), got " +
                      std::to_string(wires.size()));
    else
        QSV_CHECK(!wires.empty(), "gate " + name + " needs at least one wire");
    QSV_CHECK((int)params.size() >= e->info.n_params,
              "gate " + name + " needs " + std::to_string(e->info.n_params) + " parameter(s)");
    check_distinct(wires);

    uint64_t ctrl = 0;
    for (int i = 0; i < e->n_ctrl; ++i) ctrl |= 1ull << bit_of(n, wires[i]);
    std::vector<int> tb;
    for (size_t i = e->n_ctrl; i < wires.size(); ++i) tb.push_back(bit_of(n, wires[i]));
    const double sgn = adjoint ? -1.0 : 1.0;
    const double t = params.empty() ? 0.0 : params[0];

    switch (e->core) {
    case C_ID: return LoweredGate{};
    case C_Z: return make_phase(ctrl | 1ull << tb[0], -1.0);
    case C_S: return make_phase(ctrl | 1ull << tb[0], adjoint ? -I1 : I1);
    case C_T: return make_phase(ctrl | 1ull << tb[0], std::polar(1.0, sgn * M_PI / 4));
    case C_PHASE: return make_phase(ctrl | 1ull << tb[0], std::polar(1.0, sgn * t));
    case C_RZ: return make_diag(tb, ctrl, {std::polar(1.0, -sgn * t / 2), std::polar(1.0, sgn * t / 2)});
    case C_IZZ:
    case C_MRZ: {
        uint64_t z = 0;
        for (int b : tb) z |= 1ull << b;
        return make_parity(z, ctrl, std::polar(1.0, -sgn * t / 2), std::polar(1.0, sgn * t / 2));
    }
    case C_SWAP: return make_two_level(tb, ctrl, 1, 2, PX);
    case C_SE: {
        Mat m = m_ry(t);  // [[c,-s],[s,c]] on span{|01>, |10>}
        return make_two_level(tb, ctrl, 1, 2, adjoint ? dagger(m, 2) : m);
    }
    case C_DE: {
        Mat m = m_ry(t);  // on span{|0011>, |1100>}
        return make_two_level(tb, ctrl, 3, 12, adjoint ? dagger(m, 2) : m);
    }
    default: {
        int dim = 1 << tb.size();
        Mat m = core_matrix(e->core, params, (int)tb.size());
        if (adjoint) m = dagger(m, dim);
        return make_dense(tb, ctrl, std::move(m));
    }
    }
}

LoweredGate lower_matrix(int n, const cplx *matrix, const std::vector<int> &ctrl_wires,
                         const std::vector<int> &tgt_wires, bool adjoint) {
    QSV_CHECK(!tgt_wires.empty(), "matrix gate needs at least one target wire");
    std::vector<int> all = ctrl_wires;
    all.insert(all.end(), tgt_wires.begin(), tgt_wires.end());
    check_distinct(all);
    uint64_t ctrl = 0;
    for (int w : ctrl_wires) ctrl |= 1ull << bit_of(n, w);
    std::vector<int> tb;
    for (int w : tgt_wires) tb.push_back(bit_of(n, w));
    int dim = 1 << tb.size();
    Mat m(matrix, matrix + (size_t)dim * dim);
    if (adjoint) m = dagger(m, dim);
    // a diagonal matrix never needs the dense path
    bool diag = tb.size() <= 4;
    for (int i = 0; diag && i < dim; ++i)
        for (int j = 0; j < dim; ++j)
            if (i != j && m[(size_t)i * dim + j] != cplx(0.0, 0.0)) {
                diag = false;
                break;
            }
    if (diag) {
        Mat d(dim);
        for (int i = 0; i < dim; ++i) d[i] = m[(size_t)i * dim + i];
        return make_diag(tb, ctrl, std::move(d));
    }
    return make_dense(tb, ctrl, std::move(m));
}

// ------------------------------------------------------------------------------------------------
// Generators.  P_11 = |1><1| projectors (GateGenerators.hpp:32-44) become control predicates, so
// <bra|G|ket> for PhaseShift / CR* / ControlledPhaseShift only reads the amplitudes with the
// projected bits set.
// ------------------------------------------------------------------------------------------------
LoweredGenerator lower_generator(int n, const std::string &name, const std::vector<int> &wires) {
    const Entry *e = find_entry(name);
    QSV_CHECK(e != nullptr && e->info.n_params == 1,
              "The operation is not supported using the adjoint differentiation method: " + name);
    if (e->info.n_wires != 0)
        QSV_CHECK((int)wires.size() == e->info.n_wires, "wrong number of wires for generator of " + name);
    check_distinct(wires);
    uint64_t ctrl = 0;
    for (int i = 0; i < e->n_ctrl; ++i) ctrl |= 1ull << bit_of(n, wires[i]);
    std::vector<int> tb;
    for (size_t i = e->n_ctrl; i < wires.size(); ++i) tb.push_back(bit_of(n, wires[i]));

    LoweredGenerator g;
    g.scale = -0.5;
    switch (e->core) {
    case C_RX: g.op = make_dense(tb, ctrl, PX); break;
    case C_RY: g.op = make_dense(tb, ctrl, PY); break;
    case C_RZ: g.op = make_diag(tb, ctrl, {1.0, -1.0}); break;
    case C_PHASE:
        g.op = make_phase(ctrl | 1ull << tb[0], 1.0);
        g.scale = 1.0;
        break;
    case C_IXX: g.op = make_dense(tb, ctrl, kron(PX, 2, PX, 2)); break;
    case C_IYY: g.op = make_dense(tb, ctrl, kron(PY, 2, PY, 2)); break;
    case C_IZZ:
    case C_MRZ: {
        uint64_t z = 0;
        for (int b : tb) z |= 1ull << b;
        g.op = make_parity(z, ctrl, 1.0, -1.0);
        break;
    }
    case C_SE:
    case C_SEM:
    case C_SEP: {
        // G = d (|00><00| + |11><11|) + Y-like block on {|01>,|10>} = d*1 + two-level [[-d,-i],[i,-d]]
        double d = e->core == C_SE ? 0.0 : (e->core == C_SEM ? 1.0 : -1.0);
        g.op = make_two_level(tb, ctrl, 1, 2, mat2(-d, -I1, I1, -d));
        g.extra_identity = d;
        break;
    }
    case C_DE:
    case C_DEM:
    case C_DEP: {
        double d = e->core == C_DE ? 0.0 : (e->core == C_DEM ? 1.0 : -1.0);
        g.op = make_two_level(tb, ctrl, 3, 12, mat2(-d, -I1, I1, -d));
        g.extra_identity = d;
        break;
    }
    default:
        fail("The operation is not supported using the adjoint differentiation method: " + name);
    }
    return g;
}

}  // namespace qsv
