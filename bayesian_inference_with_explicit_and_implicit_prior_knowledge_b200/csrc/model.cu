// model.cu — model descriptor (host -> device), error plumbing, version.
//
// pgas_model_create stands in for condSequentialMonteCarlo.__init__ (reference src/PGAS.py:24-43)
// plus the closure built by generate_Hilbert_BasisFunction (src/BasisFunctions.py:59-66): it
// uploads the data, the Gaussian likelihood and GP-input map parameters, and converts the
// (M, D) integer frequency table into the packed tensor-product ROW layout the kernels use
// (common.cuh: DevModel / ROW_*).
#include <stdarg.h>
#include <algorithm>
#include <map>
#include <vector>
#include "common.cuh"

static thread_local char g_err[512] = "";

void pgas_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

long long g_pgas_launches = 0;

extern "C" const char* pgas_last_error(void) { return g_err; }
extern "C" long long pgas_launch_count(void) { return __atomic_load_n(&g_pgas_launches, __ATOMIC_RELAXED); }
extern "C" int pgas_version(void) { return 200; }
extern "C" int pgas_abi_version(void) { return PGAS_ABI_VERSION; }
extern "C" int pgas_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// lower Cholesky of a small SPD matrix (n <= 4); returns false if not positive definite
static bool small_chol(const double* A, int n, int lda, double* L /* n x n, row-major, lda n */) {
    for (int i = 0; i < n * n; ++i) L[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double d = A[j * lda + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = A[i * lda + j];
            for (int k = 0; k < j; ++k) v -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = v / d;
        }
    }
    return true;
}

extern "C" int pgas_model_create(const pgas_model_params* p, pgas_model** out) {
    if (!p || !out) PGAS_FAIL(-1, "pgas_model_create: null argument");
    if (p->n_x < 1 || p->n_x > PGAS_MAX_NX) PGAS_FAIL(-2, "n_x=%d outside [1,%d]", p->n_x, PGAS_MAX_NX);
    if (p->n_y < 1 || p->n_y > PGAS_MAX_NY) PGAS_FAIL(-2, "n_y=%d outside [1,%d]", p->n_y, PGAS_MAX_NY);
    if (p->n_u < 0 || p->n_u > PGAS_MAX_NU) PGAS_FAIL(-2, "n_u=%d outside [0,%d]", p->n_u, PGAS_MAX_NU);
    if (p->D < 1 || p->D > PGAS_MAX_D) PGAS_FAIL(-2, "D=%d outside [1,%d]", p->D, PGAS_MAX_D);
    if (p->M < 1 || p->T < 2) PGAS_FAIL(-2, "need M >= 1 and T >= 2 (M=%d, T=%d)", p->M, p->T);
    if (!p->freq || !p->observations) PGAS_FAIL(-1, "freq / observations must not be null");
    if (p->n_u > 0 && !p->inputs) PGAS_FAIL(-1, "inputs must not be null when n_u > 0");
    if (p->idx_step < 1 || p->idx_start < 1) PGAS_FAIL(-2, "idx_start and idx_step must be >= 1");
    if (p->map_kind == PGAS_MAP_VEHICLE_SLIP && (p->n_x < 2 || p->n_u < 2))
        PGAS_FAIL(-2, "vehicle slip map needs n_x >= 2 and n_u >= 2");
    if (p->map_kind != PGAS_MAP_AFFINE && p->map_kind != PGAS_MAP_VEHICLE_SLIP && p->map_kind != PGAS_MAP_PROGRAM)
        PGAS_FAIL(-2, "unknown map_kind %d", p->map_kind);
    if (p->map_kind == PGAS_MAP_PROGRAM) {
        // validate the expression program: known opcodes, operands in range, stack within bounds, exactly D results
        if (p->prog_len < 1 || p->prog_len > PGAS_MAX_PROG) PGAS_FAIL(-2, "expression program of %d instructions (1..%d)", p->prog_len, PGAS_MAX_PROG);
        int sp = 0;
        for (int i = 0; i < p->prog_len; ++i) {
            const int op = p->prog_op[i] & 0xff, arg = p->prog_op[i] >> 8;
            if (op == PGAS_OP_PUSH_X) { if (arg < 0 || arg >= p->n_x) PGAS_FAIL(-2, "program instruction %d: state component %d", i, arg); ++sp; }
            else if (op == PGAS_OP_PUSH_U) { if (arg < 0 || arg >= p->n_u) PGAS_FAIL(-2, "program instruction %d: input component %d", i, arg); ++sp; }
            else if (op == PGAS_OP_PUSH_C) { if (arg < 0 || arg >= PGAS_MAX_PROG) PGAS_FAIL(-2, "program instruction %d: constant %d", i, arg); ++sp; }
            else if ((op >= PGAS_OP_ADD && op <= PGAS_OP_DIV) || op == PGAS_OP_POW || op == PGAS_OP_ATAN2) { if (sp < 2) PGAS_FAIL(-2, "program instruction %d: stack underflow", i); --sp; }
            else if (op >= PGAS_OP_NEG && op <= PGAS_OP_ABS) { if (sp < 1) PGAS_FAIL(-2, "program instruction %d: stack underflow", i); }
            else PGAS_FAIL(-2, "program instruction %d: unknown opcode %d", i, op);
            if (sp > PGAS_PROG_STACK) PGAS_FAIL(-2, "program instruction %d: more than %d operands on the stack", i, PGAS_PROG_STACK);
        }
        if (sp != p->D) PGAS_FAIL(-2, "expression program leaves %d values, the basis has D = %d inputs", sp, p->D);
    }

    const int D = p->D, M = p->M;
    DevModel dm;
    memset(&dm, 0, sizeof(dm));
    dm.n_x = p->n_x; dm.n_y = p->n_y; dm.n_u = p->n_u; dm.D = D; dm.M = M; dm.T = p->T;
    dm.f_start = p->idx_start; dm.f_step = p->idx_step;
    dm.map_kind = p->map_kind; dm.flags = p->flags;
    dm.norm = 1.0;
    for (int d = 0; d < D; ++d) {
        if (!(p->half_width[d] > 0.0)) PGAS_FAIL(-2, "half_width[%d] must be positive", d);
        dm.center[d] = p->center[d];
        dm.L[d] = p->half_width[d];
        dm.inv2L[d] = 1.0 / (2.0 * p->half_width[d]);
        dm.norm *= sqrt(1.0 / p->half_width[d]);          // src/BasisFunctions.py:78-79
        dm.bz[d] = p->bz[d];
        for (int k = 0; k < PGAS_MAX_NX + PGAS_MAX_NU; ++k) dm.Az[d][k] = 0.0;
        // Az is given over [state (n_x); input (n_u)]; the kernels index inputs at column n_x + k
        for (int k = 0; k < p->n_x + p->n_u; ++k) dm.Az[d][k] = p->Az[d][k];
    }
    dm.slip_lf = p->slip_lf; dm.slip_lr = p->slip_lr;
    dm.prog_len = (p->map_kind == PGAS_MAP_PROGRAM) ? p->prog_len : 0;
    for (int i = 0; i < dm.prog_len; ++i) { dm.prog_op[i] = p->prog_op[i]; dm.prog_const[i] = p->prog_const[i]; }
    if (p->lik_prog_len != 0) {
        // likelihood program: validated like the map program (operands in range, stack within bounds, ONE result) and appended
        // to the same instruction / constant arrays; its PUSH_C arguments are rebased onto the shared constant pool
        if (p->lik_prog_len < 0 || dm.prog_len + p->lik_prog_len > PGAS_MAX_PROG)
            PGAS_FAIL(-2, "likelihood program of %d instructions next to a map program of %d (limit %d together)", p->lik_prog_len, dm.prog_len, PGAS_MAX_PROG);
        int ncst = 0, sp = 0;
        for (int i = 0; i < dm.prog_len; ++i) if ((dm.prog_op[i] & 0xff) == PGAS_OP_PUSH_C) ncst = std::max(ncst, (dm.prog_op[i] >> 8) + 1);
        int nlc = 0;
        for (int i = 0; i < p->lik_prog_len; ++i) {
            const int op = p->lik_prog_op[i] & 0xff, arg = p->lik_prog_op[i] >> 8;
            if (op == PGAS_OP_PUSH_X) { if (arg < 0 || arg >= p->n_x) PGAS_FAIL(-2, "likelihood program instruction %d: state component %d", i, arg); ++sp; }
            else if (op == PGAS_OP_PUSH_U) { if (arg < 0 || arg >= p->n_u) PGAS_FAIL(-2, "likelihood program instruction %d: input component %d", i, arg); ++sp; }
            else if (op == PGAS_OP_PUSH_Y) { if (arg < 0 || arg >= p->n_y) PGAS_FAIL(-2, "likelihood program instruction %d: observation component %d", i, arg); ++sp; }
            else if (op == PGAS_OP_PUSH_C) { if (arg < 0 || arg >= PGAS_MAX_PROG) PGAS_FAIL(-2, "likelihood program instruction %d: constant %d", i, arg); nlc = std::max(nlc, arg + 1); ++sp; }
            else if ((op >= PGAS_OP_ADD && op <= PGAS_OP_DIV) || op == PGAS_OP_POW || op == PGAS_OP_ATAN2) { if (sp < 2) PGAS_FAIL(-2, "likelihood program instruction %d: stack underflow", i); --sp; }
            else if (op >= PGAS_OP_NEG && op <= PGAS_OP_ABS) { if (sp < 1) PGAS_FAIL(-2, "likelihood program instruction %d: stack underflow", i); }
            else PGAS_FAIL(-2, "likelihood program instruction %d: unknown opcode %d", i, op);
            if (sp > PGAS_PROG_STACK) PGAS_FAIL(-2, "likelihood program instruction %d: more than %d operands on the stack", i, PGAS_PROG_STACK);
        }
        if (sp != 1) PGAS_FAIL(-2, "likelihood program leaves %d values, expected the log-density alone", sp);
        if (ncst + nlc > PGAS_MAX_PROG) PGAS_FAIL(-2, "map and likelihood programs use %d constants together (limit %d)", ncst + nlc, PGAS_MAX_PROG);
        dm.lik_off = dm.prog_len; dm.lik_len = p->lik_prog_len; dm.lik_coff = ncst;
        for (int i = 0; i < p->lik_prog_len; ++i) dm.prog_op[dm.lik_off + i] = p->lik_prog_op[i];
        for (int i = 0; i < nlc; ++i) dm.prog_const[ncst + i] = p->lik_prog_const[i];
    }
    for (int r = 0; r < p->n_y; ++r) {
        dm.h0[r] = p->h0[r];
        for (int k = 0; k < p->n_x; ++k) dm.H[r][k] = p->H[r][k];
    }
    {   // R = Lr Lr^T; Rw = Lr^-1; constant of the log-density
        double Lr[PGAS_MAX_NY * PGAS_MAX_NY], Rm[PGAS_MAX_NY * PGAS_MAX_NY];
        const int n = p->n_y;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Rm[i * n + j] = p->R[i][j];
        if (!small_chol(Rm, n, n, Lr)) PGAS_FAIL(-3, "observation covariance R is not positive definite");
        double logdet = 0.0;
        for (int j = 0; j < n; ++j) {
            logdet += log(Lr[j * n + j]);
            dm.Rw[j][j] = 1.0 / Lr[j * n + j];
            for (int i = j + 1; i < n; ++i) {
                double v = 0.0;
                for (int k = j; k < i; ++k) v -= Lr[i * n + k] * dm.Rw[k][j];
                dm.Rw[i][j] = v / Lr[i * n + i];
            }
        }
        dm.R_logc = -0.5 * n * log(2.0 * M_PI) - logdet;
    }
    {
        double Lp[PGAS_MAX_NX * PGAS_MAX_NX], Pm[PGAS_MAX_NX * PGAS_MAX_NX];
        const int n = p->n_x;
        for (int i = 0; i < n; ++i) { dm.m0[i] = p->m0[i]; for (int j = 0; j < n; ++j) Pm[i * n + j] = p->P0[i][j]; }
        if (!small_chol(Pm, n, n, Lp)) PGAS_FAIL(-3, "initial covariance P0 is not positive definite");
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) dm.P0c[i][j] = Lp[i * n + j];
    }

    // ---- packed row layout --------------------------------------------------------------
    // lattice positions pos = (freq - idx_start) / idx_step; leading dims 0..D-2, last dim D-1
    std::vector<int> pos((size_t)M * D);
    for (int mI = 0; mI < M; ++mI)
        for (int d = 0; d < D; ++d) {
            const int f = p->freq[(size_t)mI * D + d] - p->idx_start;
            if (f < 0 || f % p->idx_step) PGAS_FAIL(-4, "frequency %d of basis %d is not on the lattice %d + k*%d",
                                                    p->freq[(size_t)mI * D + d], mI, p->idx_start, p->idx_step);
            pos[(size_t)mI * D + d] = f / p->idx_step;
            dm.npos = std::max(dm.npos, f / p->idx_step + 1);
        }
    // rows = distinct leading-position tuples, sorted by decreasing last-dimension extent so that
    // the non-zero tiles of every position step are a prefix of the column tiles
    std::map<std::vector<int>, int> row_of;
    std::vector<std::vector<int>> row_lead;
    std::vector<int> row_len, m_row(M);
    int kp = 0;
    for (int d = 0; d < PGAS_MAX_D; ++d) dm.npos_d[d] = 0;
    for (int mI = 0; mI < M; ++mI) {
        std::vector<int> lead(pos.begin() + (size_t)mI * D, pos.begin() + (size_t)mI * D + (D - 1));
        auto it = row_of.find(lead);
        if (it == row_of.end()) {
            it = row_of.insert({lead, (int)row_lead.size()}).first;
            row_lead.push_back(lead);
            row_len.push_back(0);
        }
        m_row[mI] = it->second;
        row_len[it->second] = std::max(row_len[it->second], pos[(size_t)mI * D + D - 1] + 1);
        kp = std::max(kp, pos[(size_t)mI * D + D - 1] + 1);
        for (int d = 0; d < D; ++d) dm.npos_d[d] = std::max(dm.npos_d[d], pos[(size_t)mI * D + d] + 1);
    }
    const int nx = p->n_x;
    dm.R = (int)row_lead.size();
    std::vector<int> order(dm.R), rank(dm.R);
    for (int r = 0; r < dm.R; ++r) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return row_len[a] > row_len[b]; });
    for (int r = 0; r < dm.R; ++r) rank[order[r]] = r;
    dm.KS = (kp + 3) / 4;
    if (dm.KS > 16) PGAS_FAIL(-20, "basis needs %d lattice positions in its last dimension; this build supports <= 64", kp);
    dm.jmax = dm.KS * 4;
    dm.NTN = (dm.R * nx + 7) / 8;
    constexpr int NTB_HOST = 5;              // keep equal to NTB in basis_eval.cuh
    dm.NTNP = (dm.NTN + NTB_HOST - 1) / NTB_HOST * NTB_HOST;
    dm.n_packed = dm.KS * dm.NTNP * 32;
    std::vector<int> perm((size_t)dm.n_packed, -1);
    for (int ks = 0; ks < 16; ++ks) dm.ntcount[ks] = 0;
    for (int mI = 0; mI < M; ++mI) {
        const int j = pos[(size_t)mI * D + D - 1], r = rank[m_row[mI]];
        for (int k = 0; k < nx; ++k) {
            const int col = r * nx + k, ks = j / 4, nt = col / 8;
            const int lane = (col % 8) * 4 + (j % 4);             // lane%4 = j%4, lane/4 = col%8
            int& slot = perm[((size_t)ks * dm.NTNP + nt) * 32 + lane];
            if (slot != -1) PGAS_FAIL(-4, "duplicate basis function index tuple (basis %d)", mI);
            slot = mI * 4 + k;
            dm.ntcount[ks] = std::max(dm.ntcount[ks], nt + 1);
        }
    }
    const int n_rows_padded = (8 * dm.NTNP + nx - 1) / nx + 1;
    std::vector<int> row_pos((size_t)n_rows_padded * MAX_LEAD, 0);
    for (int r = 0; r < dm.R; ++r)
        for (int d = 0; d < D - 1; ++d) row_pos[(size_t)rank[r] * MAX_LEAD + d] = row_lead[r][d];

    // ---- row-walk layout (two- and three-dimensional bases): [slice][block][j][i][k] ----------------------
    // rows = positions of dimension D-2 in blocks of RW_RB, walked positions j = dimension D-1; D = 3: one slice of blocks per
    // position of dimension 0 (D = 2: a single slice)
    std::vector<int> rw_perm;
    dm.rw_ok = 0; dm.rw_nblk = 0; dm.rw_slots = 0; dm.rw_nslice = 0;
    for (int b = 0; b < RW_MAXBLK; ++b) dm.rw_blen[b] = 0;
    for (int sl = 0; sl < RW_MAXSLICE; ++sl) { dm.rw_slice_nblk[sl] = 0; dm.rw_slice_off[sl] = 0; dm.rw_slice_blk[sl] = 0; }
    if (D == 2 || D == 3) {
        const int dr = D - 2, dj = D - 1;                                    // row / walked dimension
        const int nslice = (D == 3) ? dm.npos_d[0] : 1;
        const int nb_per = (dm.npos_d[dr] + RW_RB - 1) / RW_RB;              // upper bound of blocks per slice
        bool fits = nslice <= RW_MAXSLICE;
        int jtop = 0;
        for (int mI = 0; mI < M; ++mI) jtop = std::max(jtop, pos[(size_t)mI * D + dj] + 1);
        // act[slice][block][j] = 1 + (largest row of the block with a selected entry at walked position j' >= j)
        std::vector<std::vector<std::vector<int>>> act(nslice, std::vector<std::vector<int>>(nb_per, std::vector<int>(jtop + 1, 0)));
        for (int mI = 0; mI < M; ++mI) {
            const int sl = (D == 3) ? pos[(size_t)mI * D] : 0, r = pos[(size_t)mI * D + dr], j = pos[(size_t)mI * D + dj];
            act[sl][r / RW_RB][j] = std::max(act[sl][r / RW_RB][j], r % RW_RB + 1);
        }
        // blocks in walk order (slice-major; empty trailing blocks of a slice are dropped), their offsets and per-position offsets
        std::vector<int> blk_off, blk_slice, blk_index;
        std::vector<std::vector<int>> joff;
        int total = 0, nblk = 0;
        for (int sl = 0; sl < nslice && fits; ++sl) {
            int used = 0;
            for (int b = 0; b < nb_per; ++b) {
                for (int j = jtop - 1; j >= 0; --j) act[sl][b][j] = std::max(act[sl][b][j], act[sl][b][j + 1]);
                if (act[sl][b][0] > 0) used = b + 1;
            }
            dm.rw_slice_nblk[sl] = used;
            dm.rw_slice_off[sl] = total;
            dm.rw_slice_blk[sl] = nblk;
            for (int b = 0; b < used; ++b) {
                if (nblk >= RW_MAXBLK) { fits = false; break; }
                act[sl][b][0] = RW_RB;       // the first position of a block carries all RW_RB rows (zeros for rows the lattice lacks): the
                                             // walk starts its accumulators with products there instead of zeroing them (rowwalk_slice)
                int cnt[RW_RB + 1] = {0, 0, 0, 0, 0};
                std::vector<int> jo(jtop + 1, 0);
                for (int j = 0; j < jtop; ++j) {
                    cnt[act[sl][b][j]]++;
                    jo[j + 1] = jo[j] + act[sl][b][j] * nx;
                }
                for (int r = 1; r <= RW_RB; ++r) fits = fits && cnt[r] <= 255;
                dm.rw_blen[nblk] = cnt[4] | (cnt[3] << 8) | (cnt[2] << 16) | (cnt[1] << 24);     // walked in this order
                blk_off.push_back(total); blk_slice.push_back(sl); blk_index.push_back(b);
                joff.push_back(jo);
                total += jo[jtop];
                ++nblk;
            }
        }
        if (fits && nblk > 0) {
            // block id of (slice, block-in-slice)
            std::vector<std::vector<int>> bid(nslice, std::vector<int>(nb_per, -1));
            for (int q = 0; q < nblk; ++q) bid[blk_slice[q]][blk_index[q]] = q;
            rw_perm.assign((size_t)total, -1);
            for (int mI = 0; mI < M; ++mI) {
                const int sl = (D == 3) ? pos[(size_t)mI * D] : 0, r = pos[(size_t)mI * D + dr], j = pos[(size_t)mI * D + dj];
                const int q = bid[sl][r / RW_RB], i = r % RW_RB;
                for (int k = 0; k < nx; ++k) rw_perm[(size_t)blk_off[q] + joff[q][j] + (size_t)i * nx + k] = mI * 4 + k;
            }
            dm.rw_ok = 1; dm.rw_nblk = nblk; dm.rw_slots = total; dm.rw_nslice = nslice;
        }
    }

    // ---- DMMA layout (two-dimensional bases, n_x = 2; basis_mma.cuh): per block of RW_RB rows and k-step s the 32 A-fragment
    //      entries of mma.sync.m8n8k4: lane -> row = lane / 4 = 2 * i_local + k, column = lane % 4 -> walked position 4 s + lane % 4
    std::vector<int> mma_perm;
    dm.mma_ok = 0; dm.mma_nblk = 0; dm.mma_slots = 0;
    for (int b = 0; b < RW_MAXBLK; ++b) dm.mma_ks[b] = 0;
    if (D == 2 && nx == 2) {
        const int nb = (dm.npos_d[0] + RW_RB - 1) / RW_RB;
        if (nb <= RW_MAXBLK) {
            std::vector<int> Lb(nb, 0), off(nb + 1, 0);
            for (int mI = 0; mI < M; ++mI) Lb[pos[(size_t)mI * D] / RW_RB] = std::max(Lb[pos[(size_t)mI * D] / RW_RB], pos[(size_t)mI * D + 1] + 1);
            bool fits = true;
            for (int b = 0; b < nb; ++b) {
                const int ks = (Lb[b] + 3) / 4;
                fits = fits && ks <= 255;
                dm.mma_ks[b] = (unsigned char)ks;
                off[b + 1] = off[b] + ks * 32;
            }
            if (fits) {
                mma_perm.assign((size_t)off[nb], -1);
                for (int mI = 0; mI < M; ++mI) {
                    const int i = pos[(size_t)mI * D], j = pos[(size_t)mI * D + 1];
                    const int b = i / RW_RB, il = i % RW_RB, sk = j / 4, c = j % 4;
                    for (int k = 0; k < nx; ++k) mma_perm[(size_t)off[b] + sk * 32 + (2 * il + k) * 4 + c] = mI * 4 + k;
                }
                dm.mma_ok = 1; dm.mma_nblk = nb; dm.mma_slots = off[nb];
            }
        }
    }

    // ---- one device arena ---------------------------------------------------------------
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_rows = al(sizeof(int) * row_pos.size()), b_perm = al(sizeof(int) * dm.n_packed), b_freq = al(sizeof(int) * M * D);
    const size_t b_obs = al(sizeof(double) * (size_t)p->T * p->n_y), b_in = al(sizeof(double) * (size_t)p->T * std::max(p->n_u, 1));
    const size_t b_rw = al(sizeof(int) * std::max<size_t>(rw_perm.size(), 1));
    const size_t b_mma = al(sizeof(int) * std::max<size_t>(mma_perm.size(), 1));
    // lattice order of the basis functions: the statistics kernel (suffstats.cu) walks its tiles in this order, so that the eight
    // columns of a DMMA fragment share their leading positions (broadcast loads) and are neighbours in the last dimension
    std::vector<int> lat_perm((size_t)M);
    for (int i = 0; i < M; ++i) lat_perm[i] = i;
    std::stable_sort(lat_perm.begin(), lat_perm.end(), [&](int a, int b) {
        for (int d = 0; d < D; ++d)
            if (pos[(size_t)a * D + d] != pos[(size_t)b * D + d]) return pos[(size_t)a * D + d] < pos[(size_t)b * D + d];
        return false;
    });
    const size_t b_lat = al(sizeof(int) * (size_t)M);
    const size_t total = b_rows + b_perm + b_freq + b_obs + b_in + b_rw + b_mma + b_lat;
    char* arena = nullptr;
    PGAS_CUDA(cudaMalloc((void**)&arena, total));
    size_t o = 0;
    auto up = [&](const void* src, size_t bytes, size_t slot) -> const void* {
        const void* d = arena + o;
        if (bytes) cudaMemcpy(arena + o, src, bytes, cudaMemcpyHostToDevice);
        o += slot;
        return d;
    };
    dm.row_pos = (const int*)up(row_pos.data(), sizeof(int) * row_pos.size(), al(sizeof(int) * row_pos.size()));
    dm.perm = (const int*)up(perm.data(), sizeof(int) * dm.n_packed, b_perm);
    dm.freq = (const int*)up(p->freq, sizeof(int) * M * D, b_freq);
    dm.obs = (const double*)up(p->observations, sizeof(double) * (size_t)p->T * p->n_y, b_obs);
    dm.inputs = (const double*)up(p->inputs, p->n_u ? sizeof(double) * (size_t)p->T * p->n_u : 0, b_in);
    dm.rw_perm = (const int*)up(rw_perm.data(), sizeof(int) * rw_perm.size(), b_rw);
    dm.mma_perm = (const int*)up(mma_perm.data(), sizeof(int) * mma_perm.size(), b_mma);
    dm.lat_perm = (const int*)up(lat_perm.data(), sizeof(int) * (size_t)M, b_lat);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(arena); PGAS_FAIL((int)e, "model upload failed: %s", cudaGetErrorString(e)); }

    pgas_model* mdl = new pgas_model;
    mdl->dev = dm;
    mdl->arena = arena;
    mdl->arena_bytes = total;
    *out = mdl;
    return 0;
}

extern "C" int pgas_model_destroy(pgas_model* model) {
    if (!model) return 0;
    cudaFree(model->arena);
    delete model;
    return 0;
}

extern "C" int pgas_model_jmax(const pgas_model* model) { return model ? model->dev.jmax : -1; }
int pgas_model_npos(const pgas_model* model) { return model->dev.npos; }
