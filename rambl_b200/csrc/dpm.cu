// Kernels of the Dirichlet-process strain clustering for sm_100a (see dpm.cuh for the mapping to
// /root/reference/StrainCall/NonparametricClustering.cpp and Strain.cpp).
//
// Numerics: the reference works in x87 long double; the device works in FP64.  A read's weight
// under a strain is kept as exp(loglik) and multiplied by the strain's current mass, where the
// reference forms exp(log(mass/total) + loglik); both feed std::discrete_distribution, which
// normalises its weights, so the two differ by rounding only (DESIGN.md states the tolerance
// and the parity tests check paths, assignments and abundances against the oracle).
#include "dpm.cuh"

#include <algorithm>
#include <cmath>

#include "common.hpp"
#include "dpm_dev.cuh"

namespace rambl {

namespace {

__global__ void __launch_bounds__(128) k_loglik(const StepGroup* __restrict__ groups, const int* __restrict__ I)
{
    const StepGroup g = groups[blockIdx.y];
    const int s = blockIdx.x;
    if (s >= g.S || g.m == 0 || g.mode == MODE_ASSIGN) return;
    __shared__ double lut[36];
    __shared__ double comp[6];
    const int tid = threadIdx.x;
    const int slot = I[g.slot_off + s];
    const int lab_o = I[g.lab_off + s], lab_l = I[g.lab_off + g.S + s];
    const double* sub = g.sub + (long long)slot * 36;
    if (tid < 6)
    {
        double c = 0;
        for (int j = 0; j < 6; ++j) c += sub[tid * 6 + j];
        comp[tid] = c;
    }
    __syncthreads();
    if (tid < 36) lut[tid] = log(sub[tid]) - log(comp[tid / 6]);
    __syncthreads();
    double* row = g.ll + (long long)slot * g.ll_stride;
    const char* lab = g.label_chars + lab_o;
    for (int r = tid; r < g.m; r += blockDim.x)
    {
        const int rid = I[g.rid_off + r];
        const char* rs = g.pool_chars + I[g.rid_off + g.m + r];
        const int rl = I[g.rid_off + 2 * g.m + r];
        double d;
        if (lab_l == 1)
        {
            const int a = letter_code(lab[0]);
            if (rl == 1) d = pair_loglik(lut, a, letter_code(rs[0]));
            else d = (a < 6) ? -INFINITY : NAN;  // one strain letter against a multi-letter key
        }
        else
        {
            d = 0;
            if (I[g.rid_off + 3 * g.m + r])
            {   // the read starts inside this collapsed node: align the tails (lines 364-375)
                int ii = lab_l, jj = rl;
                while (ii > 0 && jj > 0) d += pair_loglik(lut, letter_code(lab[--ii]), letter_code(rs[--jj]));
            }
            else
            {   // the read was already running: align the heads (lines 376-387)
                int ii = 0, jj = 0;
                while (ii < lab_l && jj < rl) d += pair_loglik(lut, letter_code(lab[ii++]), letter_code(rs[jj++]));
            }
        }
        // Strain::update_read_loglik; rows start at 0, so "create" == "add".  A (strain, read) pair normally occurs once
        // per level, which makes this a plain add.  A read with TWO entries on one level (graphs that are not strictly
        // levelled) is added twice in whatever order the atomics land: (ll + d1) + d2 and (ll + d2) + d1 can differ in the
        // last bit.  The device walk (walk.cu) adds such repeats in entry order, like the reference; this kernel only serves
        // subgroups the walk hands over.
        atomicAdd(&row[rid], d);
    }
}

constexpr int WEIGHTS_UNROLL = 4;

__global__ void __launch_bounds__(128) k_weights(const StepGroup* __restrict__ groups, const int* __restrict__ I,
                                                 double* __restrict__ W)
{
    const StepGroup g = groups[blockIdx.y];
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.mode == MODE_NONE || d >= g.D) return;
    const int r = I[g.draw_off + d];
    const int mate = I[g.draw_off + g.D + d];
    const int rid = I[g.rid_off + r];
    double* w = group_weights(W, g);
    // blockIdx.z strides over the strains: a single subgroup has few draws per level, and one thread walking
    // all S strains is S dependent gathers in a row
    double v[WEIGHTS_UNROLL];
    for (int s0 = blockIdx.z * WEIGHTS_UNROLL; s0 < g.S; s0 += gridDim.z * WEIGHTS_UNROLL)
    {
#pragma unroll
        for (int k = 0; k < WEIGHTS_UNROLL; ++k)
        {
            const int s = min(s0 + k, g.S - 1);
            const double* row = g.ll + (long long)I[g.slot_off + s] * g.ll_stride;
            v[k] = row[rid];
            if (mate >= 0) v[k] += row[mate];
        }
#pragma unroll
        for (int k = 0; k < WEIGHTS_UNROLL; ++k)
            if (s0 + k < g.S) w[weight_index(d, s0 + k, g.S)] = exp(v[k]);
    }
    if (g.mode == MODE_GIBBS && blockIdx.z == 0)
    {
        const int rl = I[g.rid_off + 2 * g.m + r];
        group_codes(W, g)[d] = (rl == 1) ? letter_code(g.pool_chars[I[g.rid_off + g.m + r]]) : 7;
    }
}

// hard_clustering, NonparametricClustering.cpp:17-125: every read copy is spread over the strains by
// its posterior; masses and substitution counts are summed and folded into the strain models.
// Deterministic by construction (no atomics): pass 1 stores the normaliser of every draw, pass 2
// gives each strain to one warp whose lanes stride over the draws with private accumulators and
// combine them in a fixed shuffle tree.
__global__ void __launch_bounds__(256) k_hard(const StepGroup* __restrict__ groups, const int* __restrict__ I,
                                              double* __restrict__ Dar, double* __restrict__ W)
{
    const StepGroup g = groups[blockIdx.x];
    if (g.mode != MODE_HARD) return;
    __shared__ double ab[DPM_SMAX];
    const int tid = threadIdx.x, S = g.S, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const unsigned full = 0xffffffffu;
    for (int s = tid; s < S; s += blockDim.x) ab[s] = Dar[g.ab_off + s];
    __syncthreads();
    double* T = group_norms(W, g);
    const double* wt = group_weights(W, g);
    for (int d = tid; d < g.D; d += blockDim.x)
    {
        double t = 0;
        for (int s = 0; s < S; ++s) t += ab[s] * wt[weight_index(d, s, S)];
        T[d] = t;
    }
    __syncthreads();
    for (int s = warp; s < S; s += nwarp)
    {
        double acc[37];
#pragma unroll
        for (int k = 0; k < 37; ++k) acc[k] = 0;
        const char* lab = g.label_chars + I[g.lab_off + s];
        const int lab_l = I[g.lab_off + S + s];
        const int la = (lab_l == 1) ? letter_code(lab[0]) : 7;
        const double a_s = ab[s];
        for (int d = lane; d < g.D; d += 32)
        {
            const double p = a_s * wt[weight_index(d, s, S)] / T[d];
            acc[0] += p;
            const int r = I[g.draw_off + d];
            const char* rs = g.pool_chars + I[g.rid_off + g.m + r];
            const int rl = I[g.rid_off + 2 * g.m + r];
            if (rl == 1)
            {
                const int b = letter_code(rs[0]);
                if (la < 6 && b < 6) acc[1 + la * 6 + b] += p;
            }
            else if (I[g.rid_off + 3 * g.m + r])
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
        double* sub = g.sub + (long long)I[g.slot_off + s] * 36;
        for (int k = 0; k < 37; ++k)
        {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
            if (lane == 0)
            {
                if (k == 0) Dar[g.ab_off + s] = v;
                else if (v != 0) sub[k - 1] += v;
            }
        }
    }
}

// np_bayes_clustering (NonparametricClustering.cpp:127-244) and read_assign (776-836): a sequential
// Gibbs chain -- draw t picks strain c with probability mass[c] * weight[c][read], then mass[c] += 1.
// The draw is std::discrete_distribution's: first strain whose cumulative weight reaches u * total
// (lower_bound over the normalised partial sums, bits/random.tcc), u from the shared
// std::mt19937(1234) stream the reference restarts on every call, and no random number at all when
// there are fewer than two strains.
//
// One CTA of GIBBS_NW warps per subgroup, 32 consecutive draws per round, draw j on lane j of EVERY
// warp, evaluated SPECULATIVELY and then corrected to the exact sequential result.  With m[] the masses
// at the start of the round, lane j's cumulative weight at strain s is  base_j(s) + corr_j(s):
//     base_j(s) = sum_{s' <= s} m[s'] * w_j[s']                      (pass 1, once per round)
//     corr_j(s) = sum_{i < j, pick_i <= s} w_j[pick_i]               (every earlier lane adds 1 to its pick)
// Pass 1 picks with corr = 0 (binary search of u * total).  Each following pass lets every lane CHECK
// its pick against the two cumulative weights around it -- 31 gathered weights instead of a walk over
// all S strains -- re-deriving it only when the check fails.  Lane 0 is exact after pass 1 and lane j
// is exact once lanes < j are, so the fixed point IS the sequential chain; masses move by 1 in
// thousands, so a round almost always settles in two passes (the counters report it).
// The warps split the two long loops of a round -- warp q owns strain chunk q of the prefix sums
// and source-lane chunk q of the gather -- and exchange partial sums through shared memory; all
// warps then hold identical picks, so the cheap steps are simply done by each of them.  The sums are
// therefore defined chunk-wise: base(s) = off[chunk] + (fma chain inside the chunk), corr = p0+p1+..,
// each p in lane order; the same definition serves the check and the re-derivation, and every
// cumulative weight stays non-decreasing in s.
// The weights of a round are one contiguous S x 32 tile, bulk-copied (TMA, no tensor map) into a
// double-buffered shared tile while the previous round is being settled.
constexpr int GIBBS_L = 32 / GIBBS_NW;  // source lanes per warp in the gather

// A wide launch (NG > 1) takes NG blocks of 32 draws per round, block b on warps 4b..4b+3.  Inside a block
// nothing changes.  Across blocks the same speculation applies one level up: the draws of block b see every
// pick of the blocks before it as a per-strain count h_b[s] (one byte per block, packed in a word per strain),
//     corr_j(s) += sum_{s' <= s} h_b[s'] * w_j[s']
// evaluated over the warp's strain chunk, where only a handful of strains have a non-zero count.  All blocks
// check at once against the picks currently published; draw 32b+j is final once every earlier draw is, so the
// fixed point is still the sequential chain.  Masses move by 1 in hundreds to tens of thousands, so a round of
// 128 draws settles in two passes most of the time as well.  A round's weights are NG consecutive tiles: still
// one bulk copy.  Shared memory grows with NG, so the launcher picks NG by strain count and by how many
// subgroups share the machine (a batch of hundreds fills the SMs with narrow CTAs instead).
__device__ __forceinline__ void group_bar(int grp)
{
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}

template <int NG>
__global__ void __launch_bounds__(128 * NG, 1)
k_gibbs(const StepGroup* __restrict__ groups, const int* __restrict__ I, double* __restrict__ Dar,
        double* __restrict__ W, const double* __restrict__ U, unsigned long long* counters, int smem_S)
{
    const StepGroup g = groups[blockIdx.x];
    if (g.mode != MODE_GIBBS && g.mode != MODE_ASSIGN) return;
    extern __shared__ __align__(128) double gibbs_smem[];  // sized for the largest S of the launch (smem_S)
    double* wbuf = gibbs_smem;                       // [2][block][strain][lane] weights of a round (bulk-copied)
    double* cumbuf = wbuf + 2 * NG * smem_S * 32;    // [block][strain][lane] prefix sums inside the strain chunk
    double* mass = cumbuf + NG * smem_S * 32;        // [smem_S] masses at the start of the round
    double* mass0 = mass + smem_S;                   // [smem_S] masses at the start of the launch
    double* ctot = mass0 + smem_S;                   // [block][GIBBS_NW][lane] chunk totals
    double* part = ctot + NG * GIBBS_NW * 32;        // [3][block][GIBBS_NW][lane] partial correction sums
    double* redo = part + 3 * NG * GIBBS_NW * 32;    // [block][GIBBS_NW][lane] the same partials while a pick is re-derived
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(redo + NG * GIBBS_NW * 32);  // [2]
    unsigned* hpack = reinterpret_cast<unsigned*>(bars + 2);  // [smem_S] picks per strain in this round, byte b = block b
    int* tcount = reinterpret_cast<int*>(hpack + smem_S);     // [smem_S] picks per strain since the launch began
    int* cnt = tcount + smem_S;                      // [smem_S][8]
    const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & (GIBBS_NW - 1), grp = tid >> 7;
    const int S = g.S, D = g.D;
    const unsigned full = 0xffffffffu;
    const double* wt = group_weights(W, g);
    const int* code = group_codes(W, g);
    const int Dp = padded_draws(D);
    for (int k = tid; k < S * 8; k += blockDim.x) cnt[k] = 0;
    for (int s = tid; s < S; s += blockDim.x)
    {
        const double a = Dar[g.ab_off + s];
        mass[s] = a; mass0[s] = a; hpack[s] = 0; tcount[s] = 0;
    }
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int tiles = Dp / 32;                       // tiles of 32 draws per sweep
    const int per_sweep = (tiles + NG - 1) / NG;     // rounds per sweep (the last one may be short of blocks)
    const long long n_rounds = (S >= 2) ? (long long)g.nsweeps * per_sweep : 0;
    unsigned long long rounds = 0, passes = 0;
    // stage the weights of a round: its tiles are contiguous, S*256 bytes each -- one bulk copy
    auto stage = [&](long long r, int blk) {
        if (tid == 0)
        {
            const int b = (int)(r & 1);
            const unsigned bytes = (unsigned)min(NG, tiles - blk * NG) * (unsigned)S * 256u;
            mbar_expect_tx(&bars[b], bytes);
            bulk_g2s(wbuf + (size_t)b * NG * smem_S * 32, wt + (long long)blk * NG * S * 32, bytes, &bars[b]);
        }
    };
    if (n_rounds > 0) stage(0, 0);
    int sweep = 0, blk = 0;
    const int Cs = (S + GIBBS_NW - 1) / GIBBS_NW;               // strains per chunk
    const int s_lo = min(S, warp * Cs), s_hi = min(S, s_lo + Cs);  // this warp's chunk
    int chunk_step = 1;
    while (chunk_step * 2 <= Cs) chunk_step *= 2;
    const unsigned below = (NG > 1) ? ((1u << (8 * grp)) - 1u) : 0u;  // the bytes of hpack that count for this block
    unsigned char* hbytes = reinterpret_cast<unsigned char*>(hpack);
    // the uniform and the read letter of a draw come from global memory: fetch them one round ahead
    double u_next = 0.0;
    int cd_next = 0;
    if (n_rounds > 0)
    {
        const int d0 = grp * 32 + lane;
        if (grp < tiles && d0 < D)
        {
            u_next = U[d0];
            if (g.mode == MODE_GIBBS) cd_next = code[d0];
        }
    }
    for (long long r = 0; r < n_rounds; ++r)
    {
        const bool active = grp < tiles - blk * NG;  // a short last round leaves the high blocks idle
        const int d = (blk * NG + grp) * 32 + lane;
        const bool valid = active && d < D;
        const double u = u_next;
        const int cd = cd_next;
        const int blk_next = (blk + 1 == per_sweep) ? 0 : blk + 1;
        if (r + 1 < n_rounds)
        {
            stage(r + 1, blk_next);  // overlaps this round's arithmetic
            const int dn = (blk_next * NG + grp) * 32 + lane;
            const int sweep_n = sweep + (blk_next == 0 ? 1 : 0);
            const bool vn = grp < tiles - blk_next * NG && dn < D;
            u_next = vn ? U[(long long)sweep_n * D + dn] : 0.0;
            cd_next = (vn && g.mode == MODE_GIBBS) ? code[dn] : 0;
        }
        mbar_wait(&bars[r & 1], (unsigned)((r >> 1) & 1));
        const double* wl = wbuf + ((size_t)(r & 1) * NG * smem_S + (size_t)grp * S) * 32 + lane;
        double* cl = cumbuf + (size_t)grp * smem_S * 32 + lane;
        double* ct = ctot + grp * GIBBS_NW * 32 + lane;
        double* pb = part + grp * GIBBS_NW * 32 + lane;  // + (which * NG * GIBBS_NW + warp) * 32
        double* rd = redo + grp * GIBBS_NW * 32 + lane;  // + warp * 32
        double off[GIBBS_NW + 1];
        off[0] = 0;
        // base(s) = offset of the chunk of s + prefix sum inside the chunk
        auto base = [&](int s) {
            double o = 0;
#pragma unroll
            for (int q = 1; q < GIBBS_NW; ++q) o = (s >= q * Cs) ? off[q] : o;
            return o + cl[s * 32];
        };
        double base_tot = 0;
        int c = -1;
        if (active)
        {
            // ---- pass 1: prefix sums of mass * weight over this warp's strain chunk, in strain order
            {
                double cum = 0;
                int s = s_lo;
                for (; s + 4 <= s_hi; s += 4)
                {
                    double m4[4], w4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { m4[q] = mass[s + q]; w4[q] = wl[(s + q) * 32]; }
#pragma unroll
                    for (int q = 0; q < 4; ++q) { cum = fma(m4[q], w4[q], cum); cl[(s + q) * 32] = cum; }
                }
                for (; s < s_hi; ++s) { cum = fma(mass[s], wl[s * 32], cum); cl[s * 32] = cum; }
                ct[warp * 32] = cum;
            }
            group_bar(grp);
#pragma unroll
            for (int q = 0; q < GIBBS_NW; ++q) off[q + 1] = off[q] + ct[q * 32];
            base_tot = off[GIBBS_NW];
            // lower_bound of u*total over the (non-decreasing) cumulative weights of strains 0..S-2, S-1 if none
            // reaches it -- in two levels: the chunk first (off[q] IS the cumulative weight at the end of chunk
            // q-1, the same number, so this is the same answer), then a binary search inside that chunk only
            const double thr = u * base_tot;
            int q = 0;
#pragma unroll
            for (int k = 1; k < GIBBS_NW; ++k) q += (off[k] < thr) ? 1 : 0;
            double oq = 0;
#pragma unroll
            for (int k = 1; k < GIBBS_NW; ++k) oq = (q == k) ? off[k] : oq;
            const int lo = q * Cs, hi = min(lo + Cs, S - 1);  // candidates lo..hi, strain S-1 only as the fallback
            int cn = lo;
            for (int step = chunk_step; step > 0; step >>= 1)
            {
                const int p = cn + step;
                if (p <= hi && oq + cl[(p - 1) * 32] < thr) cn = p;
            }
            c = valid ? min(cn, S - 1) : -1;
        }
        ++passes;
        // ---- settle: check every pick against the picks of the earlier draws until nothing moves
        int settle_passes = 0;
        int c_pub = -1;  // what this lane has in hpack (warp 0 of a block keeps it)
        for (;;)
        {
            if (NG > 1)
            {
                if (active && warp == 0)
                {
                    if (c_pub >= 0) hbytes[c_pub * 4 + grp] = 0;
                    __syncwarp();
                    const unsigned same = __match_any_sync(full, c);
                    if (c >= 0 && lane == __ffs(same) - 1) hbytes[c * 4 + grp] = (unsigned char)__popc(same);
                    c_pub = c;
                }
                __syncthreads();
            }
            bool moved = false;
            if (active)
            {
                // this warp gathers source lanes [warp*L, warp*L + L) of its block: picks, then weights, then sums
                int ci[GIBBS_L];
                double wi[GIBBS_L];
#pragma unroll
                for (int i = 0; i < GIBBS_L; ++i) ci[i] = __shfl_sync(full, c, warp * GIBBS_L + i);
#pragma unroll
                for (int i = 0; i < GIBBS_L; ++i) wi[i] = wl[max(ci[i], 0) * 32];
                double p_prev = 0, p_here = 0, p_tot = 0;
#pragma unroll
                for (int i = 0; i < GIBBS_L; ++i)
                {
                    // a term counts with multiplier 1.0 or 0.0: fma(w, 1, acc) == acc + w and fma(w, 0, acc) == acc
                    const bool live = (warp * GIBBS_L + i < lane) && (ci[i] >= 0);
                    p_tot = fma(wi[i], live ? 1.0 : 0.0, p_tot);
                    p_here = fma(wi[i], (live && ci[i] <= c) ? 1.0 : 0.0, p_here);
                    p_prev = fma(wi[i], (live && ci[i] < c) ? 1.0 : 0.0, p_prev);
                }
                if (NG > 1 && grp > 0)
                {   // the picks of the earlier blocks, as counts over this warp's strain chunk (uniform branch)
                    for (int s = s_lo; s < s_hi; ++s)
                    {
                        const unsigned hp = hpack[s] & below;
                        if (hp)
                        {
                            const double h = (double)__dp4a(hp, 0x01010101u, 0u);
                            const double w = wl[s * 32];
                            p_tot = fma(w, h, p_tot);
                            p_here = fma(w, (s <= c) ? h : 0.0, p_here);
                            p_prev = fma(w, (s < c) ? h : 0.0, p_prev);
                        }
                    }
                }
                pb[(0 * NG * GIBBS_NW + warp) * 32] = p_tot;
                pb[(1 * NG * GIBBS_NW + warp) * 32] = p_here;
                pb[(2 * NG * GIBBS_NW + warp) * 32] = p_prev;
                group_bar(grp);
                double tot = 0, le_here = 0, le_prev = 0;  // corr(S-1), corr(c), corr(c-1): chunk partials added in order
#pragma unroll
                for (int q = 0; q < GIBBS_NW; ++q)
                {
                    tot += pb[(0 * NG * GIBBS_NW + q) * 32];
                    le_here += pb[(1 * NG * GIBBS_NW + q) * 32];
                    le_prev += pb[(2 * NG * GIBBS_NW + q) * 32];
                }
                const double thr = u * (base_tot + tot);
                bool ok = true;
                if (valid)
                {
                    const bool lo_ok = (c == 0) || (base(c - 1) + le_prev < thr);
                    const bool hi_ok = (c == S - 1) || !(base(c) + le_here < thr);
                    ok = lo_ok && hi_ok;
                }
                // A pick that fails its check is re-derived by the whole block, 32 candidate strains at a time
                // (candidate s on lane s%32 of every warp), with the SAME chunk-wise sums the check uses: warp q
                // adds its source lanes and then its strain chunk, the four partials are added in order.  The four
                // warps hold identical values, so they agree on who failed and take the barriers together.
                unsigned failed = __ballot_sync(full, !ok);
                while (failed)
                {
                    const int f = __ffs(failed) - 1;
                    failed &= failed - 1;
                    const double thr_f = __shfl_sync(full, thr, f);
                    double off_f[GIBBS_NW];
#pragma unroll
                    for (int q = 1; q < GIBBS_NW; ++q) off_f[q] = __shfl_sync(full, off[q], f);
                    const double* wf = wl - lane + f;  // the weights of draw f
                    double wv[GIBBS_L];
#pragma unroll
                    for (int i = 0; i < GIBBS_L; ++i) wv[i] = wf[max(ci[i], 0) * 32];
                    int cn = S - 1;
                    for (int s0 = 0; s0 < S - 1; s0 += 32)
                    {
                        const int s = s0 + lane;
                        double p = 0;
#pragma unroll
                        for (int i = 0; i < GIBBS_L; ++i)
                        {
                            const bool live = (warp * GIBBS_L + i < f) && (ci[i] >= 0) && (ci[i] <= s);
                            p = fma(wv[i], live ? 1.0 : 0.0, p);
                        }
                        if (NG > 1 && grp > 0)
                        {
                            for (int t = s_lo; t < s_hi; ++t)
                            {
                                const unsigned hp = hpack[t] & below;
                                if (hp) p = fma(wf[t * 32], (t <= s) ? (double)__dp4a(hp, 0x01010101u, 0u) : 0.0, p);
                            }
                        }
                        rd[warp * 32] = p;
                        group_bar(grp);
                        double corr = 0;
#pragma unroll
                        for (int q = 0; q < GIBBS_NW; ++q) corr += rd[q * 32];
                        double o = 0;
#pragma unroll
                        for (int q = 1; q < GIBBS_NW; ++q) o = (s >= q * Cs) ? off_f[q] : o;
                        // strain S-1 is the fallback of the lower bound: it and the padding lanes count as "reached"
                        const bool reached = (s >= S - 1) || !((o + cl[min(s, S - 1) * 32 - lane + f]) + corr < thr_f);
                        const unsigned hit = __ballot_sync(full, reached);
                        group_bar(grp);  // rd is rewritten by the next window
                        if (hit) { cn = min(s0 + __ffs(hit) - 1, S - 1); break; }
                    }
                    if (lane == f)
                    {
                        moved = moved || (cn != c);
                        c = cn;
                    }
                }
            }
            ++passes;
            // draw 32b+j is final after 32b+j+1 passes, so 32*NG+1 passes always suffice -- the cap only guards
            // the device against a launch that does not terminate
            const int any_moved = __syncthreads_or(moved ? 1 : 0);  // also orders this pass before the next publication
            if (!any_moved) break;
            if (++settle_passes > 32 * NG + 8)
            {   // cannot happen for finite weights; report it instead of committing a round that is not the chain
                if (tid == 0 && counters) atomicAdd(&counters[2], 1ull);
                break;
            }
        }
        ++rounds;
        // ---- commit: letter statistics, then the masses of the next round from the exact pick counts
        if (valid && warp == 0 && g.mode == MODE_GIBBS) atomicAdd(&cnt[c * 8 + cd], 1);
        if (NG > 1)
        {
            for (int s = tid; s < S; s += blockDim.x)
            {
                const unsigned hp = hpack[s];
                if (hp)
                {
                    const int tc = tcount[s] + (int)__dp4a(hp, 0x01010101u, 0u);
                    tcount[s] = tc;
                    mass[s] = mass0[s] + (double)tc;
                    hpack[s] = 0;
                }
            }
        }
        else
        {
            if (valid && warp == 0) atomicAdd(&tcount[c], 1);
            __syncthreads();
            for (int s = tid; s < S; s += blockDim.x) mass[s] = mass0[s] + (double)tcount[s];
        }
        __syncthreads();
        if (blk_next == 0) ++sweep;
        blk = blk_next;
    }
    if (tid == 0 && counters)
    {
        atomicAdd(&counters[0], rounds);
        atomicAdd(&counters[1], passes);
    }
    if (tid >= 32) return;
    // normalise the masses; fold the averaged counts into the models (lines 217-243)
    double z = 0;
    for (int s = 0; s < S; ++s) z += mass[s];
    for (int s = lane; s < S; s += 32)
    {
        double v = mass[s] / z;
        if (g.mode == MODE_GIBBS) v *= (double)g.read_size;
        Dar[g.ab_off + s] = v;
        if (g.mode == MODE_GIBBS && I[g.lab_off + S + s] == 1)
        {
            const int la = letter_code(g.label_chars[I[g.lab_off + s]]);
            if (la < 6)
            {
                double* sub = g.sub + (long long)I[g.slot_off + s] * 36 + la * 6;
                for (int b = 0; b < 6; ++b)
                    if (cnt[s * 8 + b]) sub[b] += (double)cnt[s * 8 + b] / (double)g.nsweeps;
            }
        }
    }
}

// np_bayes_clustering / read_assign with ONE warp per 32-draw block: the chain itself is gibbs_w_chain (dpm_dev.cuh).
// NS = strains per lane in the per-warp bookkeeping: 2 for levels of up to 64 strains, 4 for up to 128
template <int NB, int NS>
__global__ void __launch_bounds__(32 * NB, 1)
k_gibbs_w(const StepGroup* __restrict__ groups, const int* __restrict__ I, double* __restrict__ Dar,
          double* __restrict__ W, const double* __restrict__ U, unsigned long long* counters, int smem_S)
{
    const StepGroup g = groups[blockIdx.x];
    if (g.mode != MODE_GIBBS && g.mode != MODE_ASSIGN) return;
    constexpr int GIBBS_LIST = gibbs_list_len<NS>();
    extern __shared__ __align__(128) double gibbs_smem[];  // sized for the largest S of the launch (smem_S <= 32*NS)
    GibbsShared gs;
    gs.wbuf = gibbs_smem;
    gs.wbuf_doubles = 2 * (size_t)NB * smem_S * 32;
    gs.masses = gs.wbuf + 2 * NB * smem_S * 32;
    gs.mass0 = gs.masses + NB * smem_S;
    gs.row_S = smem_S;
    gs.bars = reinterpret_cast<unsigned long long*>(gs.mass0 + smem_S);
    gs.hpacks = gs.bars + 2;
    gs.lists = reinterpret_cast<uint2*>(gs.hpacks + 2 * smem_S);
    gs.pmask = reinterpret_cast<unsigned*>(gs.lists + NB * GIBBS_LIST);
    gs.cnt = reinterpret_cast<int*>(gs.pmask + NB * smem_S);
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = g.S;
    if (tid == 0) { mbar_init(&gs.bars[0], 1); mbar_init(&gs.bars[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned uses0 = 0u, uses1 = 0u;
    unsigned long long rounds = 0, passes = 0;
    // (the chain starts with a CTA barrier, which also publishes the barrier initialisation)
    gibbs_w_chain<NB, NS, true>(gs, uses0, uses1, NB, S, g.D, g.nsweeps, g.mode == MODE_GIBBS, group_weights(W, g), group_codes(W, g), U,
                                Dar + g.ab_off, rounds, passes, counters);
    if (tid == 0 && counters)
    {
        atomicAdd(&counters[0], rounds);
        atomicAdd(&counters[1], passes);
    }
    if (tid >= 32) return;
    // normalise the masses; fold the averaged counts into the models (lines 217-243)
    const double* mass = gs.masses;
    const int* cnt = gs.cnt;
    double z = 0;
    for (int s = 0; s < S; ++s) z += mass[s];
    for (int s = lane; s < S; s += 32)
    {
        double v = mass[s] / z;
        if (g.mode == MODE_GIBBS) v *= (double)g.read_size;
        Dar[g.ab_off + s] = v;
        if (g.mode == MODE_GIBBS && I[g.lab_off + S + s] == 1)
        {
            const int la = letter_code(g.label_chars[I[g.lab_off + s]]);
            if (la < 6)
            {
                double* sub = g.sub + (long long)I[g.slot_off + s] * 36 + la * 6;
                for (int bb = 0; bb < 6; ++bb)
                    if (cnt[s * 8 + bb]) sub[bb] += (double)cnt[s * 8 + bb] / (double)g.nsweeps;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_inherit(const InheritOp* __restrict__ ops)
{
    const InheritOp op = ops[blockIdx.y];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < op.ll_stride) op.ll[(long long)op.dst * op.ll_stride + i] = op.ll[(long long)op.src * op.ll_stride + i];
    if (blockIdx.x == 0 && threadIdx.x < 36) op.sub[(long long)op.dst * 36 + threadIdx.x] = op.sub[(long long)op.src * 36 + threadIdx.x];
}

// Strain::Strain(int N, DoubleL e) with N = 100 (Strain.cpp:42-71, NonparametricClustering.cpp:281)
__global__ void k_init_models(double* sub, int n_slots, double e)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots * 36) return;
    const int q = i % 36;
    sub[i] = (q / 6 == q % 6) ? 100 * (1 - e) : 100 * e;
}

}  // namespace

// dynamic shared memory of k_gibbs<NG> (the carve-up at the top of the kernel)
static size_t gibbs_smem_bytes(int ng, int smem_S)
{
    return sizeof(double) * ((size_t)3 * ng * smem_S * 32 + 2 * (size_t)smem_S + (size_t)5 * ng * GIBBS_NW * 32 + 2) +
           sizeof(int) * ((size_t)smem_S * 10);
}

static int g_gibbs_blocks = 0;  // rambl_set_gibbs_blocks

template <int NG>
static void launch_gibbs(const StepLaunch& L, int smem_S, cudaStream_t st)
{
    const size_t smem = gibbs_smem_bytes(NG, smem_S);
    static size_t configured_on[64] = {0};  // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_on[dev & 63];
    if (smem > configured)
    {
        RAMBL_CUDA(cudaFuncSetAttribute(k_gibbs<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_gibbs<NG><<<L.n_groups, 128 * NG, smem, st>>>(L.groups, L.iarena, L.darena, L.weights, L.uniforms, L.counters, smem_S);
}

// dynamic shared memory of k_gibbs_w<NB>
static size_t gibbs_w_smem_bytes(int nb, int smem_S)
{
    const size_t list = smem_S <= 64 ? 72 : 136;  // GIBBS_LIST of the NS the launcher picks
    return 8 * ((size_t)64 * nb * smem_S + (size_t)(nb + 3) * smem_S + 2 + (size_t)nb * list) + 4 * ((size_t)nb * smem_S + 8 * (size_t)smem_S);
}

template <int NB, int NS>
static void launch_gibbs_w_ns(const StepLaunch& L, int smem_S, cudaStream_t st)
{
    const size_t smem = gibbs_w_smem_bytes(NB, smem_S);
    static size_t configured_on[64] = {0};  // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_on[dev & 63];
    if (smem > configured)
    {
        RAMBL_CUDA(cudaFuncSetAttribute(k_gibbs_w<NB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_gibbs_w<NB, NS><<<L.n_groups, 32 * NB, smem, st>>>(L.groups, L.iarena, L.darena, L.weights, L.uniforms, L.counters, smem_S);
}

// block counts the warp-per-block kernel is built for: shared memory, not a power of two, decides
static void launch_gibbs_w(int nb, const StepLaunch& L, int smem_S, cudaStream_t st)
{
    if (smem_S <= 64)
        switch (nb)
        {
            case 8: return launch_gibbs_w_ns<8, 2>(L, smem_S, st);
            case 7: return launch_gibbs_w_ns<7, 2>(L, smem_S, st);
            case 6: return launch_gibbs_w_ns<6, 2>(L, smem_S, st);
            case 5: return launch_gibbs_w_ns<5, 2>(L, smem_S, st);
            case 4: return launch_gibbs_w_ns<4, 2>(L, smem_S, st);
            case 3: return launch_gibbs_w_ns<3, 2>(L, smem_S, st);
            case 2: return launch_gibbs_w_ns<2, 2>(L, smem_S, st);
            default: return launch_gibbs_w_ns<1, 2>(L, smem_S, st);
        }
    switch (nb)
    {
        case 8: case 7: case 6: return launch_gibbs_w_ns<6, 4>(L, smem_S, st);  // 65+ strains never fit more
        case 5: return launch_gibbs_w_ns<5, 4>(L, smem_S, st);
        case 4: return launch_gibbs_w_ns<4, 4>(L, smem_S, st);
        case 3: return launch_gibbs_w_ns<3, 4>(L, smem_S, st);
        case 2: return launch_gibbs_w_ns<2, 4>(L, smem_S, st);
        default: return launch_gibbs_w_ns<1, 4>(L, smem_S, st);
    }
}

void set_gibbs_blocks(int blocks) { g_gibbs_blocks = blocks; }

void launch_level_step(const StepLaunch& L, cudaStream_t st, int* launches)
{
    if (L.n_groups == 0) return;
    if (L.max_S > DPM_SMAX) throw Error(RAMBL_ERR_CAPACITY, "more candidate strains than DPM_SMAX");
    if (L.max_m > 0 && L.max_S > 0)
    {
        k_loglik<<<dim3(L.max_S, L.n_groups), 128, 0, st>>>(L.groups, L.iarena);
        ++*launches;
    }
    if (L.max_D > 0 && (L.any_hard || L.any_gibbs))
    {
        // enough CTAs to cover the machine: a batch of many subgroups already has them, a single one spreads its strains
        const int d_blocks = (L.max_D + 127) / 128;
        const int z_max = (L.max_S + WEIGHTS_UNROLL - 1) / WEIGHTS_UNROLL;
        const int z = std::max(1, std::min(z_max, (4 * 148) / std::max(1, d_blocks * L.n_groups)));
        k_weights<<<dim3(d_blocks, L.n_groups, z), 128, 0, st>>>(L.groups, L.iarena, L.weights);
        ++*launches;
    }
    if (L.any_hard)
    {
        k_hard<<<L.n_groups, 256, 0, st>>>(L.groups, L.iarena, L.darena, L.weights);
        ++*launches;
    }
    if (L.any_gibbs)
    {
        if (L.gibbs_begin) RAMBL_CUDA(cudaEventRecord(L.gibbs_begin, st));
        const int smem_S = (L.max_S + 3) & ~3;
        // 32-draw blocks per round: a few subgroups leave SMs idle, so their chains go wide (as far as the
        // shared memory of one SM carries S strains); a batch of hundreds fills the SMs with narrow CTAs,
        // several per SM.  Levels of at most 128 strains take the warp-per-block kernel, wider ones the
        // four-warps-per-block kernel.  The chain -- and so every result -- is the same for any choice.
        int nb = 8;
        bool warp_per_block = L.max_S <= 128;
        if (g_gibbs_blocks > 0) nb = g_gibbs_blocks;
        if (g_gibbs_blocks < 0) { nb = -g_gibbs_blocks; warp_per_block = false; }
        if (warp_per_block)
        {
            // the widest CTA that still lets every subgroup of the launch be resident at once
            const size_t sm_bytes = 227 * 1024;
            while (nb > 1 && (gibbs_w_smem_bytes(nb, smem_S) > sm_bytes ||
                              (g_gibbs_blocks == 0 && 148 * (sm_bytes / gibbs_w_smem_bytes(nb, smem_S)) < (size_t)L.n_groups)))
                nb = g_gibbs_blocks > 0 ? nb >> 1 : nb - 1;  // pinned counts stay powers of two
            // a batch that leaves one block per chain and few CTAs per SM (wide levels) needs the warps of the
            // four-warps-per-block kernel to hide latency (measured on 500 subgroups: 3.5 s against 4.3 s)
            if (g_gibbs_blocks == 0 && nb == 1 && L.max_S > 64 && L.n_groups > 148) warp_per_block = false;
        }
        if (warp_per_block)
        {
            launch_gibbs_w(nb, L, smem_S, st);
        }
        else
        {
            nb = nb > 4 ? 4 : nb;
            if (g_gibbs_blocks == 0 && L.n_groups > 148) nb = 1;
            while (nb > 1 && gibbs_smem_bytes(nb, smem_S) > 227 * 1024) nb >>= 1;
            if (nb == 4) launch_gibbs<4>(L, smem_S, st);
            else if (nb == 2) launch_gibbs<2>(L, smem_S, st);
            else launch_gibbs<1>(L, smem_S, st);
        }
        if (L.gibbs_end) RAMBL_CUDA(cudaEventRecord(L.gibbs_end, st));
        ++*launches;
    }
    RAMBL_CUDA(cudaGetLastError());
}

void launch_inherit(const InheritOp* d_ops, int n_ops, long long max_stride, cudaStream_t st, int* launches)
{
    if (n_ops == 0) return;
    k_inherit<<<dim3((unsigned)((max_stride + 255) / 256), n_ops), 256, 0, st>>>(d_ops);
    ++*launches;
    RAMBL_CUDA(cudaGetLastError());
}

void launch_init_models(double* sub, int n_slots, double e, cudaStream_t st, int* launches)
{
    k_init_models<<<(n_slots * 36 + 255) / 256, 256, 0, st>>>(sub, n_slots, e);
    ++*launches;
    RAMBL_CUDA(cudaGetLastError());
}

}  // namespace rambl
