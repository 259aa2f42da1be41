// The device-resident strain walk for sm_100a (see walk.cuh for what it replaces and why).
//
// One CTA of NB warps -- or a cluster of such CTAs -- owns one subgroup from "^" to "$".  Every phase of a level is a loop
// over the threads of the CTA(s) with barriers in between; the Gibbs sweeps are gibbs_w_chain (dpm_dev.cuh), the same code
// the level-synchronous kernel k_gibbs_w runs, with up to NB 32-draw blocks per CTA and round; the bookkeeping between levels (pruning, path extension,
// the 80-candidate cut, slot assignment) is done by warp 0 with ballots and scans, in the order the reference's loops
// visit candidates and edges, because strain order, tie-breaking and the floating-point sums depend on that order.
// Everything a level reads from the host side is static (uploaded once); nothing goes back until the walk ends.
#include "walk.cuh"

#include <algorithm>
#include <cstring>

#include "common.hpp"
#include "dpm_dev.cuh"

namespace rambl {

namespace {

constexpr unsigned long long HASH_MUL = 1099511628211ull;        // FNV-1a, as engine.cpp keys strain sequences
constexpr unsigned long long LEN_MUL = 0x9e3779b97f4a7c15ull;

struct WalkShared  // carved out of dynamic shared memory after the Gibbs arrays
{
    unsigned long long* keys;  // [SMAX]
    double* ab;             // [SMAX] candidate abundances
    double* al;             // [SMAX] increments of this level
    double* delta;          // [SMAX]
    int* slot;              // [SMAX]
    int* lab_off;           // [SMAX]
    int* lab_len;           // [SMAX]
    int* idx;               // [SMAX] kept candidates, in order
    int* la;                // [SMAX] letter code of a one-letter strain label, 8 for a collapsed node
    unsigned char* parent_used;  // [SMAX]
    int* v;                 // scalars shared by the CTA (V_*)
};
enum { V_NCAND = 0, V_CB, V_BRANCH, V_TRAIL, V_FREE, V_NOPS, V_STATUS, V_D, V_COUNT };

__device__ __forceinline__ unsigned long long extend_hash(unsigned long long h, const char* s, int n)
{
    for (int i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= HASH_MUL; }
    return h;
}

template <typename T>
__device__ __forceinline__ T* carve(unsigned char*& p, size_t n)
{
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 15) & ~size_t(15);
    return r;
}

}  // namespace

size_t walk_smem_bytes(int nb, int tile_S, bool cluster)
{
    auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t b = 0;
    b += al(sizeof(double) * 2 * (size_t)nb * tile_S * 32);      // wbuf
    b += al(sizeof(double) * (size_t)nb * WALK_SMAX);            // masses
    b += al(sizeof(double) * WALK_SMAX);                         // mass0
    b += al(sizeof(unsigned long long) * 2);                     // bars
    b += al(sizeof(unsigned long long) * 2 * WALK_SMAX);         // hpacks
    b += al(sizeof(uint2) * (size_t)nb * (32 * 4 + 8));          // lists (NS = 4)
    b += al(sizeof(unsigned) * (size_t)nb * WALK_SMAX);          // pmask
    b += al(sizeof(int) * WALK_SMAX * 8);                        // cnt
    b += al(sizeof(unsigned long long) * WALK_SMAX);
    b += 3 * al(sizeof(double) * WALK_SMAX);
    b += 5 * al(sizeof(int) * WALK_SMAX);
    b += al(WALK_SMAX);
    b += al(sizeof(int) * 16);
    if (cluster) b += al(sizeof(unsigned short) * 2 * GIBBS_CMAX * WALK_SMAX) + al(sizeof(unsigned) * 2 * GIBBS_CMAX);
    return b + 128;
}

namespace {

// CL: one CLUSTER of CTAs per subgroup, for batches with fewer subgroups than the GPU has SMs.  All CTAs of the cluster
// walk the levels in step: the data-parallel phases of a level (slot copies, read set, log-likelihood update, weights,
// hard_clustering) are split over the threads of the whole cluster with cluster barriers in between, every CTA adds NB
// blocks of 32 draws to a round of the Gibbs chain (gibbs_w_chain<.., CL = true>), and rank 0 alone does the
// bookkeeping between levels and tells the others what the next level looks like through a small descriptor in global
// memory (w->helper: candidates, buffer, copies, branching flag, stop code).
template <int NB, bool CL>
__global__ void __launch_bounds__(32 * NB, (NB >= 8 ? 1 : 12 / NB)) k_walk(const WalkSub* __restrict__ subs, WalkParams prm, int tile_S)
{
    constexpr int NS = 4;
    constexpr int NT = 32 * NB;
    const int crank = CL ? (int)cluster_ctarank() : 0, csize = CL ? (int)cluster_nctarank() : 1;
#define PHASE_BARRIER() do { if (CL) cluster_barrier(); else __syncthreads(); } while (0)
    const WalkSub* __restrict__ w = subs + (CL ? cluster_id_x() : blockIdx.x);
    extern __shared__ __align__(128) unsigned char walk_smem[];
    unsigned char* sp = walk_smem;
    GibbsShared gs;
    gs.wbuf = carve<double>(sp, 2 * (size_t)NB * tile_S * 32);
    gs.wbuf_doubles = 2 * (size_t)NB * tile_S * 32;
    gs.masses = carve<double>(sp, (size_t)NB * WALK_SMAX);
    gs.mass0 = carve<double>(sp, WALK_SMAX);
    gs.row_S = WALK_SMAX;
    gs.bars = carve<unsigned long long>(sp, 2);
    gs.hpacks = carve<unsigned long long>(sp, 2 * WALK_SMAX);
    gs.lists = carve<uint2>(sp, (size_t)NB * gibbs_list_len<NS>());
    gs.pmask = carve<unsigned>(sp, (size_t)NB * WALK_SMAX);
    gs.cnt = carve<int>(sp, WALK_SMAX * 8);
    if (CL)
    {
        gs.ctot = carve<unsigned short>(sp, 2 * GIBBS_CMAX * WALK_SMAX);
        gs.cflag = carve<unsigned>(sp, 2 * GIBBS_CMAX);
    }
    WalkShared ws;
    ws.keys = carve<unsigned long long>(sp, WALK_SMAX);
    ws.ab = carve<double>(sp, WALK_SMAX);
    ws.al = carve<double>(sp, WALK_SMAX);
    ws.delta = carve<double>(sp, WALK_SMAX);
    ws.slot = carve<int>(sp, WALK_SMAX);
    ws.lab_off = carve<int>(sp, WALK_SMAX);
    ws.lab_len = carve<int>(sp, WALK_SMAX);
    ws.idx = carve<int>(sp, WALK_SMAX);
    ws.la = carve<int>(sp, WALK_SMAX);
    ws.parent_used = carve<unsigned char>(sp, WALK_SMAX);
    ws.v = carve<int>(sp, 16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    // ---- immutable per-subgroup tables
    const int* const label_off = w->label_off;
    const char* const label_chars = w->label_chars;
    const int* const out_off = w->out_off;
    const int* const out_to = w->out_to;
    const int* const out_cover = w->out_cover;
    const int end_node = w->end_node;
    const int n_levels = w->n_levels;
    const int* const lvl_ent_off = w->lvl_ent_off;
    const unsigned char* const lvl_dup = w->lvl_dup;
    const unsigned* const ent_rid = w->ent_rid;
    const unsigned char* const ent_cn = w->ent_cn;
    const char* const ent_char1 = w->ent_char1;
    const int* const lvl_moff = w->lvl_moff;
    const unsigned* const m_soff = w->m_soff;
    const unsigned char* const m_len = w->m_len;
    const char* const m_chars = w->m_chars;
    const int* const pair_off = w->pair_off;
    const int* const pair_val = w->pair_val;
    const int R = w->R;
    // ---- state (plain pointers: written and re-read inside this kernel)
    double* const ll = w->ll;
    double* const sub = w->sub;
    unsigned char* const present = w->present;
    int* const free_slots = w->free_slots;
    int2* const trail = w->trail;
    double* const W = w->W;
    int* const ent_doff = w->ent_doff;
    int* const draw_entry = w->draw_entry;
    int* const draw_mate = w->draw_mate;
    unsigned char* const fresh = w->fresh;
    double* const ab_io = w->ab_io;
    int2* const ops = w->ops;
    double* const kid_ab = w->kid_ab;
    double* const lut_g = w->lut;
    WalkResult* const res = w->res;

    unsigned uses0 = 0, uses1 = 0;
    unsigned long long rounds = 0, passes = 0;
    long long n_draws = 0, n_updates = 0, n_pairs = 0, n_gbytes = 0, sum_S = 0;  // thread 0 only
    int n_glev = 0, n_unstaged = 0, max_S = 0;

    int* const helper = w->helper;
    const int gtid = crank * NT + tid, GT = csize * NT;      // this thread / all threads of the subgroup's CTA(s)
    const int gwarp = crank * NB + (tid >> 5), GW = csize * NB;
    if (tid == 0) { mbar_init(&gs.bars[0], 1); mbar_init(&gs.bars[1], 1); }
    if (crank == 0 && tid == 0)
    {
        // the root strain, Strain(100,e) (NonparametricClustering.cpp:281); level 0 extends it by "^" and sets abundance 1
        WalkCand c;
        c.slot = 0; c.node = 0; c.tail = 0; c.pad = 0; c.ab = 1.0;
        c.hash = extend_hash(1469598103934665603ull, label_chars + label_off[0], label_off[1] - label_off[0]);
        c.len = 0;
        w->cand[0][0] = c;
        trail[0] = make_int2(-1, 0);
        const int cap = w->slot_cap;
        for (int k = 1; k < cap; ++k) free_slots[k - 1] = cap - k;  // the stack hands out slots 1, 2, ..
        ws.v[V_NCAND] = 1; ws.v[V_CB] = 0; ws.v[V_BRANCH] = 0; ws.v[V_TRAIL] = 1; ws.v[V_FREE] = cap - 1;
        ws.v[V_NOPS] = 0; ws.v[V_STATUS] = -1; ws.v[V_D] = 0;
        if (CL) { helper[0] = 1; helper[1] = 0; helper[2] = 0; helper[3] = 0; helper[4] = -1; }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    int level = 0;
    for (; level < n_levels; ++level)
    {
        int S, cb, n_ops, branch;
        if (CL)
        {
            cluster_barrier();  // rank 0 has finished the bookkeeping of the previous level
            S = __ldcg(helper); cb = __ldcg(helper + 1); n_ops = __ldcg(helper + 2); branch = __ldcg(helper + 3);
            if (__ldcg(helper + 4) >= 0) break;
        }
        else { S = ws.v[V_NCAND]; cb = ws.v[V_CB]; n_ops = ws.v[V_NOPS]; branch = ws.v[V_BRANCH]; }
        const WalkCand* const cs = w->cand[cb];
        WalkCand* const nx = w->cand[cb ^ 1];
        // ---- slot copies queued by the last extension (children beyond the first take a copy of their parent)
        for (int o = 0; o < n_ops; ++o)
        {
            const int2 op = ops[o];
            const double* __restrict__ src = ll + (long long)op.x * R;  // distinct slots: the rows never overlap
            double* __restrict__ dst = ll + (long long)op.y * R;
            for (int x = gtid; x < R; x += 4 * GT)
            {
                double v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (x + k * GT < R) ? src[x + k * GT] : 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (x + k * GT < R) dst[x + k * GT] = v[k];
            }
            if (crank == 0)
                for (int q = tid; q < 36; q += NT) sub[(long long)op.y * 36 + q] = sub[(long long)op.x * 36 + q];
        }
        if (level == n_levels - 1) break;  // "$": the host closes the result (sort + merge_strains)
        if (S == 0) break;
        const int e0 = lvl_ent_off[level], m = lvl_ent_off[level + 1] - e0;
        const int mo = lvl_moff[level];  // -1: every entry of the level is one letter; else its entries' (offset, length) table
        const int mode = (m > 0) ? (branch ? MODE_GIBBS : MODE_HARD) : MODE_NONE;
        if (S > WALK_SMAX)
        {
            if (crank == 0 && tid == 0) ws.v[V_STATUS] = WALK_TOO_MANY_STRAINS;
            break;
        }
        for (int s = tid; s < S; s += NT)
        {
            const WalkCand c = cs[s];
            ws.slot[s] = c.slot;
            ws.ab[s] = c.ab;
            ws.lab_off[s] = label_off[c.node];
            ws.lab_len[s] = label_off[c.node + 1] - label_off[c.node];
        }
        PHASE_BARRIER();  // the slot copies are complete (all CTAs), the candidates are in shared memory
        int D = 0, nsweeps = 0;
        if (mode != MODE_NONE)
        {
            // ---- the level's reads and draws (NonparametricClustering.cpp:36-39,169-189,343-391)
            bool multi = false;
            for (int s = tid; s < S; s += NT) multi = multi || ws.lab_len[s] > 1;
            const int any_multi = __syncthreads_or(multi ? 1 : 0);
            if (warp == 0)
            {   // first draw of every entry: running sum of the copy numbers
                int run = 0;
                for (int r0 = 0; r0 < m; r0 += 32)
                {
                    const int r = r0 + lane;
                    const int c = r < m ? (int)ent_cn[e0 + r] : 0;
                    int x = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const int y = __shfl_up_sync(full, x, o);
                        if (lane >= o) x += y;
                    }
                    if (r < m) ent_doff[r] = run + x - c;
                    run += __shfl_sync(full, x, 31);
                }
                if (lane == 0) { ent_doff[m] = run; ws.v[V_D] = run; }
            }
            __syncthreads();
            D = ws.v[V_D];
            // "new" = first time any strain sees the read; hard_clustering only looks at the flag on collapsed nodes
            for (int r = gtid; r < m; r += GT)
            {
                const unsigned er = ent_rid[e0 + r];
                const int rid = (int)(er & 0x7fffffffu);
                const unsigned char f = ((er >> 31) || present[rid]) ? 0 : 1;  // a second entry of the read on this level: not new
                present[rid] = 1;
                fresh[r] = (mode == MODE_HARD && !any_multi) ? 0 : f;
                const int cn = (int)ent_cn[e0 + r], d0 = ent_doff[r], po = pair_off[rid];
                for (int k = 0; k < cn; ++k)
                {   // copies are counted down: the first draw of a read looks up the mate of its last copy
                    draw_entry[d0 + k] = r;
                    draw_mate[d0 + k] = pair_val[po + cn - k - 1];
                }
            }
            PHASE_BARRIER();
            if (mode == MODE_GIBBS)
            {   // a mate no strain has seen yet does not count
                for (int d = gtid; d < D; d += GT)
                {
                    const int mate = draw_mate[d];
                    if (mate >= 0 && !present[mate]) draw_mate[d] = -1;
                }
            }
            else
            {   // Strain::logprob(uid) creates the mate's entry
                PHASE_BARRIER();
                for (int d = gtid; d < D; d += GT)
                {
                    const int mate = draw_mate[d];
                    if (mate >= 0) present[mate] = 1;
                }
            }
            nsweeps = (mode == MODE_GIBBS) ? min(prm.n, 40000 / max(D, 1)) : 0;
            // ---- log-likelihood update: ll[strain][read] += log p(read letters | strain letters)
            // the strains' log tables first (Strain::logprob(a,b) = log(sub[a,b]) - log(comp[a]), Strain.cpp:130-133) ...
            for (int q = gtid; q < S * 36; q += GT)
            {
                const int st = q / 36, k = q - st * 36;
                const double* sb = sub + (long long)ws.slot[st] * 36;
                double c = 0;
                for (int j = 0; j < 6; ++j) c += sb[(k / 6) * 6 + j];
                lut_g[q] = log(sb[k]) - log(c);
            }
            for (int st = tid; st < S; st += NT) ws.la[st] = (ws.lab_len[st] == 1) ? letter_code(label_chars[ws.lab_off[st]]) : 8;
            PHASE_BARRIER();
            // ... then one thread per read-pool entry, four strains in flight (independent rows): a read has one entry per
            // level, except the entries flagged as repeats, which are added afterwards in entry order
            auto entry_term = [&](int st, int r, const char* rs, int rl) -> double {
                const double* lut = lut_g + st * 36;
                const int la = ws.la[st];
                if (la < 8)
                {
                    if (rl == 1) return pair_loglik(lut, la, letter_code(rs[0]));
                    return (la < 6) ? -INFINITY : NAN;  // one strain letter against a multi-letter key
                }
                const char* lab = label_chars + ws.lab_off[st];
                const int lab_l = ws.lab_len[st];
                double d = 0;
                if (fresh[r])
                {   // the read starts inside this collapsed node: align the tails (lines 364-375)
                    int ii = lab_l, jj = rl;
                    while (ii > 0 && jj > 0) d += pair_loglik(lut, letter_code(lab[--ii]), letter_code(rs[--jj]));
                }
                else
                {   // the read was already running: align the heads (lines 376-387)
                    int ii = 0, jj = 0;
                    while (ii < lab_l && jj < rl) d += pair_loglik(lut, letter_code(lab[ii++]), letter_code(rs[jj++]));
                }
                return d;
            };
            // (pass 0: the first entry of every read; pass 1: second entries -- distinct reads again, so still one thread per
            // entry -- after a barrier; a read with three or more entries on a level takes the ordered loop below)
            const int dup_depth = lvl_dup[level];
            for (int pass = 0; pass < (dup_depth == 1 ? 2 : 1); ++pass)
            {
                if (pass) PHASE_BARRIER();
                for (int r = gtid; r < m; r += GT)
                {
                    const unsigned er = ent_rid[e0 + r];
                    if ((int)(er >> 31) != pass) continue;
                    const int rid = (int)(er & 0x7fffffffu);
                    const char* rs = mo >= 0 ? m_chars + m_soff[mo + r] : ent_char1 + e0 + r;
                    const int rl = mo >= 0 ? (int)m_len[mo + r] : 1;
                    for (int s0 = 0; s0 < S; s0 += 4)
                    {
                        double dv[4], old[4];
                        double* pr[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const int st = min(s0 + k, S - 1);
                            dv[k] = entry_term(st, r, rs, rl);
                            pr[k] = ll + (long long)ws.slot[st] * R + rid;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) old[k] = *pr[k];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (s0 + k < S) *pr[k] = old[k] + dv[k];
                    }
                }
            }
            if (dup_depth > 1)
            {
                PHASE_BARRIER();
                for (int st = gtid; st < S; st += GT)
                {
                    double* row = ll + (long long)ws.slot[st] * R;
                    for (int r = 0; r < m; ++r)
                    {
                        const unsigned er = ent_rid[e0 + r];
                        if (!(er >> 31)) continue;
                        const char* rs = mo >= 0 ? m_chars + m_soff[mo + r] : ent_char1 + e0 + r;
                        row[er & 0x7fffffffu] += entry_term(st, r, rs, mo >= 0 ? (int)m_len[mo + r] : 1);
                    }
                }
            }
            PHASE_BARRIER();
            // ---- weights: exp(loglik(read) + loglik(mate)) per (draw, strain), tile-major (dpm_dev.cuh)
            double* const wt = W;
            double* const norms = W + (long long)S * padded_draws(D);
            int* const codes = reinterpret_cast<int*>(norms + D);
            for (int d = gtid; d < D; d += GT)
            {
                const int r = draw_entry[d];
                const int rid = (int)(ent_rid[e0 + r] & 0x7fffffffu);
                const int mate = draw_mate[d];
                for (int s0 = 0; s0 < S; s0 += 4)
                {
                    double v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const int s = min(s0 + k, S - 1);
                        const double* row = ll + (long long)ws.slot[s] * R;
                        v[k] = row[rid];
                        if (mate >= 0) v[k] += row[mate];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (s0 + k < S) wt[weight_index(d, s0 + k, S)] = exp(v[k]);
                }
                if (mode == MODE_GIBBS)
                {
                    const int rl = mo >= 0 ? (int)m_len[mo + r] : 1;
                    codes[d] = (rl == 1) ? letter_code(mo >= 0 ? m_chars[m_soff[mo + r]] : ent_char1[e0 + r]) : 7;
                }
            }
            // the Gibbs chain reads its tiles through the async proxy (bulk copies): order the generic-proxy stores
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence_block();
            PHASE_BARRIER();
            if (mode == MODE_HARD)
            {
                // ---- hard_clustering: soft assignment, masses and substitution counts (deterministic: no atomics)
                for (int d = gtid; d < D; d += GT)
                {
                    double t = 0;
                    for (int s = 0; s < S; ++s) t += ws.ab[s] * wt[weight_index(d, s, S)];
                    norms[d] = t;
                }
                PHASE_BARRIER();
                for (int s = gwarp; s < S; s += GW)
                {
                    double acc[37];
#pragma unroll
                    for (int k = 0; k < 37; ++k) acc[k] = 0;
                    const char* lab = label_chars + ws.lab_off[s];
                    const int lab_l = ws.lab_len[s];
                    const int la = (lab_l == 1) ? letter_code(lab[0]) : 7;
                    const double a_s = ws.ab[s];
                    for (int d = lane; d < D; d += 32)
                    {
                        const double p = a_s * wt[weight_index(d, s, S)] / norms[d];
                        acc[0] += p;
                        const int r = draw_entry[d];
                        const char* rs = mo >= 0 ? m_chars + m_soff[mo + r] : ent_char1 + e0 + r;
                        const int rl = mo >= 0 ? (int)m_len[mo + r] : 1;
                        if (rl == 1)
                        {
                            const int b = letter_code(rs[0]);
                            if (la < 6 && b < 6) acc[1 + la * 6 + b] += p;
                        }
                        else if (fresh[r])
                        {
                            int ii = lab_l, jj = rl;
                            while (ii > 0 && jj > 0)
                            {
                                const int a = letter_code(lab[--ii]), b = letter_code(rs[--jj]);
                                if (a < 6 && b < 6) acc[1 + a * 6 + b] += p;
                            }
                        }
                        else
                        {
                            int ii = 0, jj = 0;
                            while (ii < lab_l && jj < rl)
                            {
                                const int a = letter_code(lab[ii++]), b = letter_code(rs[jj++]);
                                if (a < 6 && b < 6) acc[1 + a * 6 + b] += p;
                            }
                        }
                    }
                    double* sb = sub + (long long)ws.slot[s] * 36;
                    for (int k = 0; k < 37; ++k)
                    {
                        double v = acc[k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
                        if (lane == 0)
                        {
                            if (k == 0) ab_io[s] = v;
                            else if (v != 0) sb[k - 1] += v;
                        }
                    }
                }
                PHASE_BARRIER();  // the increments of every strain, whichever CTA's warp summed them
                for (int s = tid; s < S; s += NT) ws.al[s] = CL ? __ldcg(ab_io + s) : ab_io[s];
            }
            else
            {
                // ---- np_bayes_clustering: the sequential Gibbs chain, NB blocks of 32 draws per round
                // as many 32-draw blocks per round as the level's tiles leave room for in the tile buffers; a level too wide
                // for even one block reads its weights from L1/L2 with all NB blocks
                const int nb_two = (int)min((size_t)NB, gs.wbuf_doubles / (2 * (size_t)S * 32));
                const int nb_one = (int)min((size_t)NB, gs.wbuf_doubles / ((size_t)S * 32));
                // a wide level: one tile buffer instead of two when that buys at least half as many blocks again
                const bool single = prm.single_buffer && nb_two < NB && 2 * nb_one >= 3 * max(nb_two, 1);
                // ... and when the weights of the whole level fit, they are copied once and stay (all NB blocks per round)
                const bool resident = (prm.single_buffer & 2) && (size_t)padded_draws(D) * S <= gs.wbuf_doubles;
                int nb = resident ? NB : (single ? nb_one : nb_two);
                if (nb == 0) nb = NB;
                gibbs_w_chain<NB, NS, false, CL>(gs, uses0, uses1, nb, S, D, nsweeps, true, wt, codes, prm.uniforms, ws.ab, rounds, passes,
                                                 prm.counters, crank, csize, resident ? 2 : (single ? 1 : 0));
                if (warp == 0)
                {   // normalise the masses; fold the averaged letter counts into the models (lines 217-243)
                    const double* mass = gs.masses;
                    double z = 0;
                    for (int s = 0; s < S; ++s) z += mass[s];
                    for (int s = lane; s < S; s += 32)
                    {
                        ws.al[s] = mass[s] / z * (double)D;
                        if (crank == 0 && ws.lab_len[s] == 1)
                        {
                            const int la = letter_code(label_chars[ws.lab_off[s]]);
                            if (la < 6)
                            {
                                double* sb = sub + (long long)ws.slot[s] * 36 + la * 6;
                                for (int bb = 0; bb < 6; ++bb)
                                    if (gs.cnt[s * 8 + bb]) sb[bb] += (double)gs.cnt[s * 8 + bb] / (double)nsweeps;
                            }
                        }
                    }
                }
            }
            if (crank == 0 && tid == 0)
            {
                n_updates += (long long)m * S;
                n_pairs += (long long)D * S;
                if (mode == MODE_GIBBS && S >= 2)
                {
                    n_draws += (long long)D * nsweeps;
                    n_gbytes += (long long)nsweeps * D * (S + 1) * 8;
                    n_glev += 1;
                    sum_S += S;
                    max_S = max(max_S, S);
                    if ((size_t)S * 32 > gs.wbuf_doubles) n_unstaged += 1;
                }
            }
            __syncthreads();
        }

        // ---- between levels (warp 0): abundances, pruning, path extension, the 80-candidate cut, slots
        if (crank == 0 && warp == 0)
        {
            int n_keep = 0;
            int free_top = ws.v[V_FREE];
            int fail = -1;
            if (mode == MODE_GIBBS)
            {
                // abundance maps keyed by the strain sequence (lines 404-429): equal sequences share an entry, the
                // last one written wins
                for (int i = lane; i < S; i += 32) ws.keys[i] = cs[i].hash ^ (cs[i].len * LEN_MUL);
                __syncwarp();
                double dmax = 0;
                for (int i = lane; i < S; i += 32)
                {
                    int j = i;
                    const unsigned long long key = ws.keys[i];
                    for (int q = S - 1; q > i; --q)
                        if (ws.keys[q] == key) { j = q; break; }
                    const double before = ws.ab[j];
                    const double after = before + ws.al[j];
                    const double delta = after - before;
                    ws.delta[i] = delta;
                    if (dmax < delta) dmax = delta;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double y = __shfl_xor_sync(full, dmax, o);
                    if (dmax < y) dmax = y;
                }
                double Z = 0;
                if (lane == 0)
                    for (int i = 0; i < S; ++i) Z += ws.al[i];
                Z = __shfl_sync(full, Z, 0);
                const double Zt = Z * prm.tau, thr = 0.01 * dmax;
                __syncwarp();
                for (int i0 = 0; i0 < S; i0 += 32)
                {
                    const int i = i0 + lane;
                    const bool in = i < S;
                    const bool keep = in && !(ws.al[i] < Zt || ws.delta[i] < thr);
                    const unsigned kb = __ballot_sync(full, keep), db = __ballot_sync(full, in && !keep);
                    if (keep) ws.idx[n_keep + __popc(kb & lt)] = i;
                    if (in && !keep) free_slots[free_top + __popc(db & lt)] = cs[i].slot;
                    if (in) ws.ab[i] += ws.al[i];
                    n_keep += __popc(kb);
                    free_top += __popc(db);
                }
            }
            else
            {
                for (int i = lane; i < S; i += 32)
                {
                    if (mode == MODE_HARD) ws.ab[i] += ws.al[i];
                    ws.idx[i] = i;
                }
                n_keep = S;
            }
            __syncwarp();
            // candidate strains of the next level (lines 473-551), in candidate order, then edge order
            int K = 0, trail_n = ws.v[V_TRAIL];
            bool branching = false;
            for (int c0 = 0; c0 < n_keep && fail < 0; c0 += 32)
            {
                const int ci = c0 + lane;
                const bool in = ci < n_keep;
                WalkCand par;
                int ea = 0, eb = 0, kids = 0, dd = 0;
                double oz = 0, moc = 0, pab = 0;
                if (in)
                {
                    par = cs[ws.idx[ci]];
                    pab = ws.ab[ws.idx[ci]];
                    ea = out_off[par.node]; eb = out_off[par.node + 1];
                    for (int e = ea; e < eb; ++e) { const double oc = (double)out_cover[e]; oz += oc; if (moc < oc) moc = oc; }
                    for (int e = ea; e < eb; ++e)
                    {
                        const double oc = (double)out_cover[e];
                        if (out_to[e] != end_node && oz > 0 && oc <= 1. && oc < moc) { dd += 1; continue; }
                        kids += 1;
                    }
                }
                branching = branching || __any_sync(full, in && (eb - ea > 1 + dd));
                int x = kids;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const int y = __shfl_up_sync(full, x, o);
                    if (lane >= o) x += y;
                }
                const int total = __shfl_sync(full, x, 31);
                if (K + total > WALK_KMAX) { fail = WALK_TOO_MANY_CHILDREN; break; }
                if (trail_n + K + total > w->trail_cap) { fail = WALK_TRAIL_FULL; break; }
                int pos = K + x - kids;
                if (in)
                    for (int e = ea; e < eb; ++e)
                    {
                        const int o = out_to[e];
                        const double oc = (double)out_cover[e];
                        WalkCand k;
                        k.ab = pab;
                        if (o != end_node && oz > 0)
                        {
                            if (oc <= 1. && oc < moc) continue;
                            k.ab = (oc > 0) ? pab * oc / oz : oz * fmin(0.01, prm.tau);
                        }
                        const int lo = label_off[o], ln = label_off[o + 1] - lo;
                        k.slot = -1;
                        k.node = o;
                        k.tail = trail_n + pos;
                        k.pad = ci;  // parent, as an index into the kept candidates
                        k.hash = extend_hash(par.hash, label_chars + lo, ln);
                        k.len = par.len + (unsigned long long)ln;
                        trail[trail_n + pos] = make_int2(par.tail, o);
                        nx[pos] = k;
                        ++pos;
                    }
                K += total;
            }
            trail_n += K;
            __syncwarp();
            if (fail < 0 && K > 80)
            {   // keep the 80 most abundant children and whatever ties with the 81st (lines 485-560)
                bool nan = false;
                for (int k = lane; k < K; k += 32) { const double a = nx[k].ab; kid_ab[k] = a; nan = nan || (a != a); }
                if (__any_sync(full, nan)) fail = WALK_NAN_IN_CUT;  // std::sort with NaN keys: the host path decides
                __syncwarp();
                if (fail < 0)
                {
                    double cut = 0;
                    for (int k0 = 0; k0 < K; k0 += 32)
                    {
                        const int k = k0 + lane;
                        const double a = k < K ? kid_ab[k] : 0;
                        int gt = 0, ge = 0;
                        for (int j = 0; j < K; ++j) { const double bq = kid_ab[j]; gt += bq > a ? 1 : 0; ge += bq >= a ? 1 : 0; }
                        const bool is_cut = k < K && gt <= 80 && 80 < ge;  // the value at index 80 of the descending sort
                        const unsigned hit = __ballot_sync(full, is_cut);
                        if (hit) { cut = __shfl_sync(full, a, __ffs(hit) - 1); break; }
                    }
                    int K2 = 0;
                    for (int k0 = 0; k0 < K; k0 += 32)
                    {
                        const int k = k0 + lane;
                        WalkCand c;
                        bool keep = false;
                        if (k < K) { c = nx[k]; keep = !(c.ab < cut); }
                        const unsigned kb = __ballot_sync(full, keep);
                        __syncwarp();
                        if (keep) nx[K2 + __popc(kb & lt)] = c;
                        K2 += __popc(kb);
                        __syncwarp();
                    }
                    K = K2;
                }
            }
            __syncwarp();
            if (fail < 0)
            {
                // slots: the first surviving child of a parent takes the parent's slot, the others copy it.  The children
                // of a parent are consecutive, so "first" is "differs from the child before"; the others pop the free stack
                // in child order (their rank among the non-first children).
                for (int i = lane; i < n_keep; i += 32) ws.parent_used[i] = 0;
                __syncwarp();
                int n_extra = 0, p_carry = -1;
                for (int k0 = 0; k0 < K; k0 += 32)
                {
                    const int k = k0 + lane;
                    const int p = k < K ? nx[k].pad : -2;
                    int p_before = __shfl_up_sync(full, p, 1);
                    if (lane == 0) p_before = p_carry;
                    const bool first = k < K && p != p_before, extra = k < K && !first;
                    const unsigned eb = __ballot_sync(full, extra);
                    if (k < K)
                    {
                        const int pslot = ws.slot[ws.idx[p]];
                        if (first) { ws.parent_used[p] = 1; nx[k].slot = pslot; }
                        else
                        {
                            const int rank = n_extra + __popc(eb & lt);
                            if (rank >= free_top) fail = WALK_OUT_OF_SLOTS;
                            else
                            {
                                const int dst = free_slots[free_top - 1 - rank];
                                ops[rank] = make_int2(pslot, dst);
                                nx[k].slot = dst;
                            }
                        }
                    }
                    n_extra += __popc(eb);
                    p_carry = __shfl_sync(full, p, 31);
                }
                if (__any_sync(full, fail >= 0)) fail = WALK_OUT_OF_SLOTS;
                __syncwarp();
                if (fail < 0)
                {
                    free_top -= n_extra;
                    for (int i0 = 0; i0 < n_keep; i0 += 32)
                    {   // parents without a surviving child give their slot back
                        const int i = i0 + lane;
                        const bool unused = i < n_keep && !ws.parent_used[i];
                        const unsigned ub = __ballot_sync(full, unused);
                        if (unused) free_slots[free_top + __popc(ub & lt)] = ws.slot[ws.idx[i]];
                        free_top += __popc(ub);
                    }
                    if (lane == 0) { ws.v[V_NOPS] = n_extra; ws.v[V_FREE] = free_top; }
                }
            }
            if (lane == 0)
            {
                ws.v[V_NCAND] = K;
                ws.v[V_CB] = cb ^ 1;
                ws.v[V_BRANCH] = branching ? 1 : 0;
                ws.v[V_TRAIL] = trail_n;
                if (fail >= 0) ws.v[V_STATUS] = fail;
                if (CL)
                {   // what the other CTAs of the cluster need to know about the next level
                    helper[0] = K; helper[1] = cb ^ 1; helper[2] = ws.v[V_NOPS]; helper[3] = branching ? 1 : 0; helper[4] = ws.v[V_STATUS];
                }
            }
        }
        __syncthreads();
        if (!CL && ws.v[V_STATUS] >= 0) break;
    }
    if (CL && crank != 0) return;  // rank 0 writes the result
    __syncthreads();
    // ---- result: the candidates at "$" with their paths, or why the walk stopped
    const int status = ws.v[V_STATUS] >= 0 ? ws.v[V_STATUS] : (level == n_levels - 1 && ws.v[V_NCAND] > 0 ? WALK_DONE : WALK_NO_CANDS);
    const int S = (status == WALK_DONE) ? ws.v[V_NCAND] : 0;
    if (status == WALK_DONE)
    {
        const WalkCand* cs = w->cand[ws.v[V_CB]];
        int* paths = w->paths;
        for (int c = tid; c < S; c += NT)
        {
            w->final_slot[c] = cs[c].slot;
            w->final_ab[c] = cs[c].ab;
            int t = cs[c].tail;
            for (int pos = n_levels - 1; pos >= 0 && t >= 0; --pos)
            {
                const int2 e = trail[t];
                paths[(long long)c * n_levels + pos] = e.y;
                t = e.x;
            }
        }
    }
    if (tid == 0)
    {
        WalkResult r;
        r.status = status;
        r.n_cands = S;
        r.levels = level;
        r.reason_level = level;
        r.draws = n_draws; r.loglik_updates = n_updates; r.weight_pairs = n_pairs; r.gibbs_bytes = n_gbytes;
        r.rounds = rounds; r.passes = passes;
        r.sum_S = sum_S; r.gibbs_levels = n_glev; r.unstaged_levels = n_unstaged; r.max_S = max_S; r.pad = 0;
        r.free_top = ws.v[V_FREE]; r.branching = ws.v[V_BRANCH]; r.cand_buf = ws.v[V_CB]; r.pad2 = 0;
        *res = r;
        if (prm.counters)
        {
            atomicAdd(&prm.counters[0], rounds);
            atomicAdd(&prm.counters[1], passes);
        }
    }
#undef PHASE_BARRIER
}

template <int NB, bool CL>
void launch_walk_nb(const WalkSub* d_subs, int n_subs, const WalkParams& prm, int tile_S, int cluster, cudaStream_t st)
{
    const size_t smem = walk_smem_bytes(NB, tile_S, CL);
    static size_t configured_on[64] = {0};  // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_on[dev & 63];
    if (smem > configured)
    {
        RAMBL_CUDA(cudaFuncSetAttribute(k_walk<NB, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (!CL)
    {
        k_walk<NB, CL><<<n_subs, 32 * NB, smem, st>>>(d_subs, prm, tile_S);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_subs * cluster));
    cfg.blockDim = dim3(32 * NB);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RAMBL_CUDA(cudaLaunchKernelEx(&cfg, k_walk<NB, CL>, d_subs, prm, tile_S));
}

}  // namespace

void launch_walk(const WalkSub* d_subs, int n_subs, const WalkParams& prm, int nb, int tile_S, int cluster, cudaStream_t st, int* launches)
{
    if (n_subs <= 0) return;
    if (cluster > 1)
    {
        if (nb != 8 || cluster > GIBBS_CMAX) throw Error(RAMBL_ERR_INVALID, "clusters need eight-warp CTAs and at most eight CTAs");
        launch_walk_nb<8, true>(d_subs, n_subs, prm, tile_S, cluster, st);
    }
    else
        switch (nb)
        {
            case 8: launch_walk_nb<8, false>(d_subs, n_subs, prm, tile_S, 1, st); break;
            case 4: launch_walk_nb<4, false>(d_subs, n_subs, prm, tile_S, 1, st); break;
            case 2: launch_walk_nb<2, false>(d_subs, n_subs, prm, tile_S, 1, st); break;
            default: launch_walk_nb<1, false>(d_subs, n_subs, prm, tile_S, 1, st); break;
        }
    ++*launches;
    RAMBL_CUDA(cudaGetLastError());
}

}  // namespace rambl
