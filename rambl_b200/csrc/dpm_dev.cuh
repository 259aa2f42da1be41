// Device-side building blocks shared by the level-synchronous kernels (dpm.cu) and the device-resident strain
// walk (walk.cu): letter codes and Strain::logprob, the layout of the weights scratch, bulk-copy (TMA without a
// tensor map) helpers, and the warp-per-block speculative Gibbs chain.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "dpm.cuh"

namespace rambl {

namespace {

constexpr int GIBBS_NW = 4;             // strain chunks of the prefix sums (and warps per 32-draw block of k_gibbs)

// A,C,G,T,-,= are the letters of the strain model (Strain.cpp:7); N is special on the strain side
// (NonparametricClustering.cpp:358,372,384); anything else is a key the model never counted.
__device__ __forceinline__ int letter_code(char c)
{
    switch (c)
    {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case '-': return 4;
        case '=': return 5;
        case 'N': return 6;
    }
    return 7;
}

// Strain::logprob(a,b) = log(sub_count[a,b]) - log(comp_count[a]) with std::map defaults for keys
// outside the 6x6 table (Strain.cpp:130-133): unknown b -> log(0); unknown a -> log(0)-log(0).
__device__ __forceinline__ double pair_loglik(const double* lut, int a, int b)
{
    if (a == 6) a = b;  // "if (ssb=="N") ssb = rrb"
    if (a < 6) return b < 6 ? lut[a * 6 + b] : -INFINITY;
    return NAN;
}

// Scratch of one group inside the weights buffer (doubles, from w_off, which is a multiple of 32):
//   [Dp/32][S][32] weights in tiles of 32 consecutive draws x S strains (Dp = D rounded up to 32): the
//           tile of a Gibbs round is ONE contiguous, 256-byte aligned run of S*256 bytes -- a single bulk
//           copy -- and inside it the 32 draws of a strain are consecutive (conflict-free lanes); the
//           padding of the last tile is never read as data;
//   [D]     normaliser per draw (k_hard);
//   [D]     the read letter of every draw as an int code (k_gibbs statistics), stored in D double slots.
__device__ __forceinline__ int padded_draws(int D) { return (D + 31) & ~31; }
__device__ __forceinline__ double* group_weights(double* W, const StepGroup& g) { return W + g.w_off; }
__device__ __forceinline__ long long weight_index(int d, int s, int S) { return ((long long)(d >> 5) * S + s) * 32 + (d & 31); }
__device__ __forceinline__ double* group_norms(double* W, const StepGroup& g)
{
    return W + g.w_off + (long long)g.S * padded_draws(g.D);
}
__device__ __forceinline__ int* group_codes(double* W, const StepGroup& g)
{
    return reinterpret_cast<int*>(W + g.w_off + (long long)g.S * padded_draws(g.D) + g.D);
}

// ---- sm_90+/sm_100 bulk asynchronous copy (TMA without a tensor map) and its transaction barrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- thread-block clusters: rank, distributed shared memory, cluster barrier
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_id_x()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
// the shared::cluster address of `p` (a shared-memory address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ unsigned dsmem_addr(const void* p, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st_u16(unsigned addr, unsigned short v)
{
    asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_u32(unsigned addr, unsigned v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void dsmem_add_u32(unsigned addr, unsigned v)
{
    asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int GIBBS_CMAX = 8;  // CTAs per cluster (the portable maximum)

// The same chain with ONE warp per 32-draw block and up to eight blocks (256 draws) per round, for levels of
// at most 128 candidate strains (the reference prunes to about 80 and usually holds 10-50).  A warp keeps its
// block to itself -- no partial sums to exchange, no barriers inside a block -- and spends about a third of
// the instructions per draw of k_gibbs, which is what bounds a round once several warps share a scheduler:
//   * cumulative weights are not stored: phase A forms the four chunk totals (four independent fma chains),
//     phase B re-runs the half chunk that holds u * total and keeps the two cumulative weights around the pick
//     in registers (the check needs nothing else), so shared memory holds the weights only;
//   * every pick of the round is published as a per-strain count (byte b of hpack = block b) and, for the
//     block's own lanes, a per-strain lane mask; draw j of block b is corrected by
//         corr_j(s) = sum_{s' <= s} (picks of s' in blocks < b  +  picks of s' on lanes < j of block b) * w_j[s']
//     summed in strain order over the strains with a non-zero count (a short per-warp list, rebuilt every
//     pass by ballot, walked four entries at a time so that the shared-memory loads overlap);
//   * a pick that fails its check is re-derived by its warp, candidate strain s on lane s % 32, with the same
//     sums in the same order, so check and re-derivation cannot disagree.
// cum(s) = off[chunk of s] + (fma chain from the start of that chunk), as in k_gibbs.
// NS = strains per lane in the per-warp bookkeeping: 2 for levels of up to 64 strains, 4 for up to 128
//
// gibbs_w_chain is the whole chain of one subgroup and level, run by a CTA of exactly NB warps, of which the first
// `nb` take a 32-draw block per round (the caller picks nb <= NB so that the round's tiles fit the tile buffers: a
// level of many strains runs fewer blocks per round, as the per-level launches of k_gibbs_w do).  The caller owns the
// shared memory (GibbsShared) and the two transaction barriers, which it initialises once per kernel; `uses` counts
// the waits done on each barrier so far, so that the phase parity survives from one call to the next (the device
// walk calls this once per graph level).  STAGED_ONLY: the weight tiles of a round always fit the shared tile buffers
// (the caller guarantees S <= tile_S); otherwise a level with S > tile_S reads its weights straight from global
// memory (L1/L2), no staging.  On return (after a CTA barrier) masses[0..S) holds the final masses and cnt the letter
// counts per strain.
//
// CL: the chain runs on a thread-block CLUSTER of `csize` CTAs (rank `crank`), one SM each -- csize * nb blocks per
// round, block rank * nb + b on warp b of CTA `rank`.  A draw is corrected by the picks of every earlier block of the
// round: the blocks of its own CTA as before, the CTAs of lower rank through their per-strain TOTALS, which every CTA
// writes into every other CTA's shared memory (distributed shared memory) once per pass, together with its verdict
// on the previous pass ("one of my picks moved"); one cluster barrier per pass.  A round is over when no CTA moved a
// pick, which every CTA learns from the same flags, so all of them leave the loop together.  Every CTA keeps its own
// copy of the masses (updated from all totals), stages its own tiles, and counts the letters of its own draws (summed
// into rank 0 at the end).
struct GibbsShared
{
    double* wbuf;                 // [2][wbuf_doubles / 2] weight tiles of a round, double-buffered (bulk-copied)
    size_t wbuf_doubles;          // capacity of both buffers together: a staged level needs 2 * nb * S * 32
    double* masses;               // [NB][row_S] masses at the start of the round, one copy per warp
    double* mass0;                // [row_S] masses at the start of the chain
    int row_S;                    // stride of the per-strain rows (>= S)
    unsigned long long* bars;     // [2] transaction barriers of the tile buffers
    unsigned long long* hpacks;   // [2][row_S] picks per strain in a round (by round parity), byte b = block b
    uint2* lists;                 // [NB][32 * NS + 8] the strains that count for block b
    unsigned* pmask;              // [NB][row_S] lanes of block b that picked s
    int* cnt;                     // [row_S][8] letter counts per strain
    unsigned short* ctot = nullptr;  // cluster only: [2][GIBBS_CMAX][row_S] picks per strain and CTA in a pass (by pass parity)
    unsigned* cflag = nullptr;       // cluster only: [2][GIBBS_CMAX] "a pick moved in my last pass"
};

template <int NS>
__host__ __device__ constexpr int gibbs_list_len() { return 32 * NS + 8; }  // every strain + padding

template <int NB, int NS, bool STAGED_ONLY, bool CL = false>
__device__ __forceinline__ void gibbs_w_chain(const GibbsShared& gs, unsigned& uses0, unsigned& uses1, int nb, int S, int D, int nsweeps,
                                              bool count_letters, const double* wt, const int* code, const double* U,
                                              const double* ab_in, unsigned long long& rounds, unsigned long long& passes,
                                              unsigned long long* counters, int crank = 0, int csize = 1, int tile_mode = 0)
{
    // tile_mode 0: two tile buffers, the next round's tiles arrive while this round settles; 1: one buffer -- a level of
    // many strains then still runs many blocks per round, the tiles of the next round are fetched when the round is over
    // (the copy is exposed, but the round is wider); 2: the weights of the WHOLE level fit the buffers -- one copy when
    // the chain starts, nothing per round (every sweep re-reads the same tiles from shared memory)
    const bool single = tile_mode == 1, resident = tile_mode == 2;
    constexpr int GIBBS_LIST = gibbs_list_len<NS>();
    if (!CL) { crank = 0; csize = 1; }
    const int gnb = csize * nb;        // blocks of 32 draws per round, over the whole cluster
    const int gb0 = crank * nb;        // the first of them that belongs to this CTA
    double* const wbuf = gs.wbuf;
    double* const masses = gs.masses;
    double* const mass0 = gs.mass0;
    unsigned long long* const bars = gs.bars;
    unsigned long long* const hpacks = gs.hpacks;
    uint2* const lists = gs.lists;
    unsigned* const pmask = gs.pmask;
    int* const cnt = gs.cnt;
    const int smem_S = gs.row_S;
    const size_t buf_doubles = (size_t)nb * S * 32;  // one round's tiles: the second buffer starts right behind
    const bool staged = STAGED_ONLY || resident || (single ? 1 : 2) * buf_doubles <= gs.wbuf_doubles;
    const int tid = threadIdx.x, lane = tid & 31, b = tid >> 5;
    const unsigned full = 0xffffffffu;
    const int Dp = padded_draws(D);
    for (int k = tid; k < S * 8; k += blockDim.x) cnt[k] = 0;
    for (int k = tid; k < NB * smem_S; k += blockDim.x) pmask[k] = 0;
    for (int s = tid; s < S; s += blockDim.x)
    {
        const double a = ab_in[s];
        mass0[s] = a; hpacks[s] = 0; hpacks[smem_S + s] = 0;
        for (int k = 0; k < NB; ++k) masses[k * smem_S + s] = a;
    }
    __syncthreads();
    // The chain is one stream of tiles: tile T = sweep * tiles + t holds draws 32t..32t+31 of that sweep, and round
    // r takes tiles r*gnb .. r*gnb+gnb-1 whatever sweep they fall in (the weights of tile t are the same in every
    // sweep), so only the very last round can be short of blocks.
    const int tiles = Dp / 32;                       // tiles of 32 draws per sweep
    const int total_tiles = (S >= 2) ? nsweeps * tiles : 0;   // <= 5000 sweeps x 40000/32 tiles
    const int n_rounds = (total_tiles + gnb - 1) / gnb;
    // rounds in which this CTA has tiles of its own: a prefix of the rounds (only the last round can be short)
    const int my_rounds = total_tiles > gb0 ? (total_tiles - gb0 + gnb - 1) / gnb : 0;
    // stage the weights of round r: this CTA's tiles are contiguous up to the end of a sweep, then wrap to tile 0
    auto stage = [&](int r) {
        if (staged && tid == 0)
        {
            const int bf = single ? 0 : (r & 1);
            const int first = r * gnb + gb0;
            int left = min(nb, total_tiles - first);
            int stage_pos = first % tiles;
            mbar_expect_tx(&bars[bf], (unsigned)left * (unsigned)S * 256u);
            double* dst = wbuf + (size_t)bf * buf_doubles;
            while (left > 0)
            {
                const int seg = min(left, tiles - stage_pos);
                bulk_g2s(dst, wt + (long long)stage_pos * S * 32, (unsigned)seg * (unsigned)S * 256u, &bars[bf]);
                dst += (size_t)seg * S * 32;
                left -= seg;
                stage_pos += seg;
                if (stage_pos == tiles) stage_pos = 0;
            }
        }
    };
    if (resident)
    {
        if (my_rounds > 0)
        {
            if (tid == 0)
            {
                const unsigned bytes = (unsigned)tiles * (unsigned)S * 256u;
                mbar_expect_tx(&bars[0], bytes);
                for (unsigned o = 0; o < bytes; o += 65536u)
                    bulk_g2s(reinterpret_cast<char*>(wbuf) + o, reinterpret_cast<const char*>(wt) + o, min(65536u, bytes - o), &bars[0]);
            }
            mbar_wait(&bars[0], uses0 & 1u);
            uses0 += 1;
        }
    }
    else if (my_rounds > 0) stage(0);
    const int Cs = (S + GIBBS_NW - 1) / GIBBS_NW;    // strains per chunk
    const unsigned long long below = (b == 0) ? 0ull : (~0ull >> (64 - 8 * b));  // the bytes of hpack that precede this block
    const unsigned below_lo = (unsigned)below, below_hi = (unsigned)(below >> 32);
    const unsigned lt = (1u << lane) - 1u;
    double* mass = masses + b * smem_S;  // this warp's copy: a round ends without a barrier (see the commit)
    int tc[NS];                          // picks of strains lane, lane+32, .. since the launch began
#pragma unroll
    for (int h = 0; h < NS; ++h) tc[h] = 0;
    int c_last = -1;                     // this lane's final pick of the previous round
    int cpass = 0, cuse = 0;             // cluster: exchanges done so far (their parity picks the buffer), buffer of the last one
    unsigned* pm = pmask + b * smem_S;
    uint2* list = lists + b * GIBBS_LIST;
    // this block's tile of the coming round: sweep and tile within the sweep; its uniform and read letter are
    // fetched from global memory one round ahead
    int sw_next = 0, t_next = gb0 + b;
    if (tiles > 0) { sw_next = t_next / tiles; t_next -= sw_next * tiles; }
    double u_next = 0.0;
    int cd_next = 0;
    if (b < nb && gb0 + b < total_tiles && t_next * 32 + lane < D)
    {
        u_next = U[(long long)sw_next * D + t_next * 32 + lane];
        if (count_letters) cd_next = code[t_next * 32 + lane];
    }
    for (int r = 0; r < n_rounds; ++r)
    {
        const bool active = b < nb && r * gnb + gb0 + b < total_tiles;  // only the last round can leave the high blocks idle
        const int t_cur = t_next;
        const int d = t_cur * 32 + lane;
        const bool valid = active && d < D;
        const double u = u_next;
        const int cd = cd_next;
        if (r + 1 < n_rounds)
        {
            if (!single && !resident && r + 1 < my_rounds) stage(r + 1);  // overlaps this round's arithmetic
            t_next += gnb;
            while (t_next >= tiles) { t_next -= tiles; ++sw_next; }
            const int dn = t_next * 32 + lane;
            const bool vn = b < nb && (r + 1) * gnb + gb0 + b < total_tiles && dn < D;
            u_next = vn ? U[(long long)sw_next * D + dn] : 0.0;
            cd_next = (vn && count_letters) ? code[dn] : 0;
        }
        if (staged && !resident && r < my_rounds)
        {
            if (single) mbar_wait(&bars[0], (uses0 + (unsigned)r) & 1u);
            else mbar_wait(&bars[r & 1], (((r & 1) ? uses1 : uses0) + (unsigned)(r >> 1)) & 1u);
        }
        unsigned long long* hpack = hpacks + (r & 1) * smem_S;
        unsigned char* hbytes = reinterpret_cast<unsigned char*>(hpack);
        const double* wl = resident ? wbuf + (size_t)t_cur * S * 32 + lane
                         : staged ? wbuf + (size_t)(single ? 0 : (r & 1)) * buf_doubles + (size_t)b * S * 32 + lane
                                  : wt + (long long)t_cur * S * 32 + lane;
        double off[GIBBS_NW + 1];
        off[0] = 0;
        double base_tot = 0, b_prev = 0, b_here = 0;  // cum(c-1) and cum(c) without corrections
        int c = -1;
        if (active)
        {
            // ---- phase A: the four chunk totals, four independent chains in strain order; each chain also leaves its
            // value half-way through the chunk (after the first `half` strains), which lets phase B start there
            double ch[GIBBS_NW], mid[GIBBS_NW];
#pragma unroll
            for (int q = 0; q < GIBBS_NW; ++q) ch[q] = 0;
            const int n_last = S - (GIBBS_NW - 1) * Cs;  // strains in the last chunk; the others are full when this is >= 0
            const int half = (Cs + 1) >> 1;
            if (n_last >= 0)
            {
                const double* mp = mass;
                const double* wp = wl;
                const int cs32 = Cs * 32;
#pragma unroll 4
                for (int k = 0; k < half; ++k, ++mp, wp += 32)
                {
                    ch[0] = fma(mp[0], wp[0], ch[0]);
                    ch[1] = fma(mp[Cs], wp[cs32], ch[1]);
                    ch[2] = fma(mp[2 * Cs], wp[2 * cs32], ch[2]);
                    if (k < n_last) ch[3] = fma(mp[3 * Cs], wp[3 * cs32], ch[3]);
                }
#pragma unroll
                for (int q = 0; q < GIBBS_NW; ++q) mid[q] = ch[q];
#pragma unroll 4
                for (int k = half; k < Cs; ++k, ++mp, wp += 32)
                {
                    ch[0] = fma(mp[0], wp[0], ch[0]);
                    ch[1] = fma(mp[Cs], wp[cs32], ch[1]);
                    ch[2] = fma(mp[2 * Cs], wp[2 * cs32], ch[2]);
                    if (k < n_last) ch[3] = fma(mp[3 * Cs], wp[3 * cs32], ch[3]);
                }
            }
            else
            {
                for (int k = 0; k < Cs; ++k)
                {
#pragma unroll
                    for (int q = 0; q < GIBBS_NW; ++q)
                    {
                        const int s = q * Cs + k;
                        if (s < S) ch[q] = fma(mass[s], wl[s * 32], ch[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < GIBBS_NW; ++q) mid[q] = 0;
            }
#pragma unroll
            for (int q = 0; q < GIBBS_NW; ++q) off[q + 1] = off[q] + ch[q];
            base_tot = off[GIBBS_NW];
            // ---- phase B: lower_bound of u*total over the cumulative weights of strains 0..S-2 (S-1 if none
            // reaches it): the chunk first, then its half, then the chain of that half once more with the comparison
            // riding along.  Cumulative weights do not decrease, so the strains below the threshold are a prefix:
            // count them and keep the chain value of the last one; cum(c) is one more step from there.
            const double thr = u * base_tot;
            int q = 0;
#pragma unroll
            for (int k = 1; k < GIBBS_NW; ++k) q += (off[k] < thr) ? 1 : 0;
            double oq = 0, mq = mid[0];
#pragma unroll
            for (int k = 1; k < GIBBS_NW; ++k) { oq = (q == k) ? off[k] : oq; mq = (q == k) ? mid[k] : mq; }
            const int lo = q * Cs, end = min(lo + Cs, S - 1);  // candidates lo..end-1
            // the second half, when every strain of the first half is a candidate below the threshold
            const bool halves = n_last >= 0 && lo + half <= end;  // else: walk the whole chunk from its start
            const bool second = halves && (oq + mq < thr);
            const int nk = halves ? half : Cs;
            int cn = second ? lo + half : lo;
            double run = second ? mq : 0.0, run_u = run;  // run_u: the chain at the last strain below the threshold
            {
                const int s0 = cn;
                const double* mp = mass + s0;
                const double* wp = wl + s0 * 32;
                for (int k = 0; k < nk; ++k)
                {
                    const bool in = s0 + k < end;
                    const int kc = in ? k : 0;
                    run = fma(mp[kc], wp[kc * 32], run);
                    const bool under = in && (oq + run < thr);
                    if (under) { run_u = run; ++cn; }
                }
            }
            b_prev = oq + run_u;  // cum(cn-1); with nothing below it, cum(lo-1) -- the same number as off[q]
            {
                const int sc = min(cn, S - 1);
                b_here = oq + fma(mass[sc], wl[sc * 32], run_u);
            }
            c = valid ? min(cn, S - 1) : -1;
        }
        ++passes;
        // ---- settle: check every pick against the picks of the earlier draws until nothing moves
        int settle_passes = 0;
        int c_pub = -1;  // what this lane has published
        int moved_cta = 1;  // the verdict of this CTA's last pass (the first exchange of a round always goes on)
        for (;;)
        {
            if (active)
            {
                if (c_pub >= 0) { hbytes[c_pub * 8 + b] = 0; pm[c_pub] = 0; }
                __syncwarp();
                const unsigned same = __match_any_sync(full, c);
                if (c >= 0 && lane == __ffs(same) - 1)
                {
                    hbytes[c * 8 + b] = (unsigned char)__popc(same);
                    pm[c] = same;
                }
                c_pub = c;
            }
            __syncthreads();
            if (CL)
            {
                // this CTA's picks per strain, and its verdict on the previous pass, into every CTA of the cluster
                unsigned short* tot_row = gs.ctot + ((size_t)(cpass & 1) * GIBBS_CMAX + crank) * smem_S;
                for (int s = tid; s < S; s += blockDim.x)
                {
                    const unsigned long long hp = hpack[s];
                    const unsigned short t = (unsigned short)__dp4a((unsigned)hp, 0x01010101u, __dp4a((unsigned)(hp >> 32), 0x01010101u, 0u));
                    for (int k = 0; k < csize; ++k) dsmem_st_u16(dsmem_addr(tot_row + s, (unsigned)k), t);
                }
                if (tid < csize) dsmem_st_u32(dsmem_addr(gs.cflag + (cpass & 1) * GIBBS_CMAX + crank, (unsigned)tid), (unsigned)moved_cta);
                cluster_barrier();
                cuse = cpass & 1;
                ++cpass;
                unsigned any = 0;
                for (int k = 0; k < csize; ++k) any |= gs.cflag[cuse * GIBBS_CMAX + k];
                if (!any) break;  // no CTA moved a pick in the last pass: what is published is the chain
            }
            const unsigned short* ctot_now = CL ? gs.ctot + (size_t)cuse * GIBBS_CMAX * smem_S : nullptr;
            bool moved = false;
            if (active)
            {
                // the strains that count for this block (picked in an earlier block or on a lane of this one), in
                // strain order, as a list of (strain, picks in earlier blocks, lanes of this block): lane i looks at
                // strains i and i+32; the list is padded to a multiple of four with entries that add nothing
                int n_list = 0;
                {
                    unsigned hs[NS], mm[NS];
#pragma unroll
                    for (int h = 0; h < NS; ++h)
                    {
                        hs[h] = 0; mm[h] = 0;
                        if (32 * h >= S) continue;  // (warp-uniform) no strain up there on this level
                        const bool in = lane + 32 * h < S;
                        const unsigned long long hp = in ? hpack[lane + 32 * h] : 0ull;
                        mm[h] = in ? pm[lane + 32 * h] : 0u;
                        hs[h] = __dp4a((unsigned)hp & below_lo, 0x01010101u, __dp4a((unsigned)(hp >> 32) & below_hi, 0x01010101u, 0u));
                        if (CL && in)
                            for (int k = 0; k < crank; ++k) hs[h] += ctot_now[k * smem_S + lane + 32 * h];  // the CTAs before this one
                    }
#pragma unroll
                    for (int h = 0; h < NS; ++h)
                    {
                        if (32 * h >= S) continue;
                        const unsigned bal = __ballot_sync(full, (hs[h] | mm[h]) != 0);
                        if (hs[h] | mm[h]) list[n_list + __popc(bal & lt)] = make_uint2((unsigned)(lane + 32 * h) | (hs[h] << 8), mm[h]);
                        n_list += __popc(bal);
                    }
                    if (lane < 4) list[n_list + lane] = make_uint2(0u, 0u);
                    __syncwarp();
                }
                // corr(S-1) and corr(c-1) as chains over the list; corr(c) is the second chain taken one term further
                // -- the term of strain c itself, which this lane looks up directly (zero if c is not on the list)
                const int cc = max(c, 0);
                const unsigned long long hpc = hpack[cc];
                const unsigned mmc = pm[cc];
                const double w_c = wl[cc * 32];
                double p_prev = 0, p_tot = 0;
                for (int i = 0; i < n_list; i += 4)
                {
                    uint2 e[4];
                    double w4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = list[i + k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) w4[k] = wl[(e[k].x & 0xffu) * 32];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const int s = (int)(e[k].x & 0xffu);
                        const double kd = (double)((int)(e[k].x >> 8) + __popc(e[k].y & lt));
                        p_tot = fma(w4[k], kd, p_tot);
                        p_prev = fma(w4[k], (s < c) ? kd : 0.0, p_prev);
                    }
                }
                int k_c = (int)__dp4a((unsigned)hpc & below_lo, 0x01010101u, __dp4a((unsigned)(hpc >> 32) & below_hi, 0x01010101u, 0u)) +
                          __popc(mmc & lt);
                if (CL)
                    for (int k = 0; k < crank; ++k) k_c += ctot_now[k * smem_S + cc];
                const double p_here = fma(w_c, (double)k_c, p_prev);
                const double thr = u * (base_tot + p_tot);
                bool ok = true;
                if (valid)
                {
                    const bool lo_ok = (c == 0) || (b_prev + p_prev < thr);
                    const bool hi_ok = (c == S - 1) || !(b_here + p_here < thr);
                    ok = lo_ok && hi_ok;
                }
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
                    const unsigned lt_f = (1u << f) - 1u;
                    int cn = S - 1;
                    double carry = 0, bp_new = 0, bh_new = 0;
                    bool hit_any = false;
                    for (int s0 = 0; s0 < S - 1 && !hit_any; s0 += 32)
                    {
                        const int s = s0 + lane;
                        const int q = (s >= Cs ? 1 : 0) + (s >= 2 * Cs ? 1 : 0) + (s >= 3 * Cs ? 1 : 0);
                        const int lo = q * Cs;
                        double acc = 0;  // chain of the chunk of s, from its start up to s
                        const int t_end = min(S, s0 + 32);
                        for (int t = 0; t < t_end; ++t)
                        {
                            const double m = mass[t], w = wf[t * 32];
                            if (t >= lo && t <= s) acc = fma(m, w, acc);
                        }
                        double o = 0;
#pragma unroll
                        for (int k = 1; k < GIBBS_NW; ++k) o = (q == k) ? off_f[k] : o;
                        const double base = o + acc;
                        double p = 0;
                        for (int i = 0; i < n_list; ++i)
                        {
                            const uint2 e = list[i];
                            const int t = (int)(e.x & 0xffu);
                            const int k = (int)(e.x >> 8) + __popc(e.y & lt_f);
                            p = fma(wf[t * 32], (t <= s) ? (double)k : 0.0, p);
                        }
                        // strain S-1 is the fallback of the lower bound: it and the padding lanes count as "reached"
                        const bool reached = (s >= S - 1) || !(base + p < thr_f);
                        const unsigned hit = __ballot_sync(full, reached);
                        if (hit)
                        {
                            const int th = __ffs(hit) - 1;
                            cn = min(s0 + th, S - 1);
                            bh_new = __shfl_sync(full, base, th);
                            const double below_hit = __shfl_sync(full, base, max(th - 1, 0));
                            bp_new = th > 0 ? below_hit : carry;
                            hit_any = true;
                        }
                        else carry = __shfl_sync(full, base, 31);
                    }
                    if (!hit_any) bp_new = carry;  // only when S-1 is a multiple of 32: the fallback strain, cum(S-2) below it
                    if (lane == f)
                    {
                        moved = moved || (cn != c);
                        c = cn; b_prev = bp_new; b_here = bh_new;
                    }
                }
            }
            ++passes;
            // draw 32b+j is final after 32b+j+1 passes, so 32*NB+1 passes always suffice -- the cap only guards
            // the device against a launch that does not terminate
            moved_cta = __syncthreads_or(moved ? 1 : 0);  // also orders this pass before the next publication
            if (!CL && !moved_cta) break;
            if (++settle_passes > 32 * NB * csize + 8)
            {   // cannot happen for finite weights; report it instead of committing a round that is not the chain
                if (tid == 0 && counters) atomicAdd(&counters[2], 1ull);
                break;
            }
        }
        ++rounds;
        if (single && r + 1 < my_rounds) stage(r + 1);  // every warp is past its last look at this round's tiles
        // ---- commit: letter statistics, then the masses of the next round from the exact pick counts
        // Every warp updates its own copy of the masses from the round's counts, so the next round starts without a
        // barrier.  The counts of round r stay readable until every warp has passed the first barrier of round
        // r+1; each warp wipes its own bytes of them in the commit of round r+1, before it publishes into that
        // buffer again in round r+2.
        if (valid && count_letters) atomicAdd(&cnt[c * 8 + cd], 1);
        if (active && c_pub >= 0) pm[c_pub] = 0;
        if (c_last >= 0) reinterpret_cast<unsigned char*>(hpacks + ((r + 1) & 1) * smem_S)[c_last * 8 + b] = 0;
        c_last = active ? c_pub : -1;
#pragma unroll
        for (int h = 0; h < NS; ++h)
        {
            const int sh = lane + 32 * h;
            if (sh < S)
            {
                const unsigned long long hp = hpack[sh];
                int picks = (int)__dp4a((unsigned)hp, 0x01010101u, __dp4a((unsigned)(hp >> 32), 0x01010101u, 0u));
                if (CL)
                    for (int k = 0; k < csize; ++k)
                        if (k != crank) picks += gs.ctot[((size_t)cuse * GIBBS_CMAX + k) * smem_S + sh];
                if (picks)
                {
                    tc[h] += picks;
                    mass[sh] = mass0[sh] + (double)tc[h];
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();  // the letter counts of every warp
    if (CL)
    {   // the letter counts of the other CTAs' draws go to rank 0
        if (crank != 0)
            for (int k = tid; k < S * 8; k += blockDim.x)
                if (cnt[k]) dsmem_add_u32(dsmem_addr(cnt + k, 0u), (unsigned)cnt[k]);
        cluster_barrier();
    }
    if (staged && !resident)
    {
        if (single) uses0 += (unsigned)my_rounds;
        else
        {
            uses0 += (unsigned)((my_rounds + 1) >> 1);
            uses1 += (unsigned)(my_rounds >> 1);
        }
    }
}

}  // namespace

}  // namespace rambl
