// Batched progressive sum-of-pairs alignment of insertion strings on sm_100a.
//
// Replaces MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char>::align
// (/root/reference/StrainCall/MultipleSequenceAlignment.hpp:87-107,
//  MultipleSequenceAlignmentSP.cpp:10-301) as PartialOrderGraph::canonize_insert_at_level calls it
// (PartialOrderGraph.cpp:449-455,507-519,546): every graph level whose insertions differ in length is
// one PROBLEM; all problems of all subgroups of a batch are solved together.
//
// The reference evaluates every DP cell with three loops over the s sequences already in the
// profile and keeps one state per (cell, sequence).  Both collapse: (1) the per-sequence state of an
// interior cell is the same for every sequence (the reference derives it from the first row of the
// profile column, MultipleSequenceAlignmentSP.cpp:208-218,235-245) and the states of the border
// cells are never "ins" where they are asked for "ins" and never asked for "del"; (2) the sum over
// sequences then only needs the letter-class counts of the column.  A cell is O(1) integer work:
// scores are integers (3,-5,-6,-2), so int32 is exact where the reference sums doubles, and the
// reference's tie order (match >= insert >= delete) is kept.
//
// Mapping.  Real insertion levels are SMALL (a handful of letters against a profile of a dozen columns, tens to
// hundreds of strings), so a warp per problem leaves most lanes idle and one oversized problem used to size the
// shared memory of every CTA.  Problems are therefore dealt into SIZE CLASSES by their longest string; a class
// gives each problem a group of G lanes (1, 8, 16 or 32) of a CTA and a shared-memory region sized for
// the class, so a CTA of the smallest class solves 128 problems at once (one lane each).  Within a group:
//   * profile columns have a stable id (creation order); `ord` maps profile position -> id, so a column
//     insertion only shifts `ord`; letters live column-major in global scratch (colchar[id][row]);
//   * the per-column sums against every letter class (what a DP cell needs) are kept INCREMENTALLY: aligning a
//     string adds one letter or one gap to every column -- 7 additions -- instead of a 7x7 recomputation;
//   * borders: the first row is a closed form, the first column a prefix sum over the group's lanes;
//   * the interior runs as an anti-diagonal wavefront over the group's lanes;
//   * the traceback is a short serial walk by one lane, but the profile update it implies is done by all lanes
//     (prefix counts over the operation string give every operation its column and its letter).
// A problem whose profile outgrows its class (more columns than the class holds) reports it and is solved again
// in the next class; the last class keeps its tables in global memory and takes anything (no compiled-in limit
// on columns or letters: the reference allocates its tables on the heap).
#include "msa_sp.hpp"

#include <algorithm>
#include <cstring>

namespace rambl {

namespace {

enum { P_MAT = 0, P_INS = 1, P_DEL = 2 };
constexpr int CLS_GAP = 4, CLS_PLUS = 5, CLS_UNK = 6, NCLS = 7;

__constant__ int c_S[NCLS][NCLS];

// SimpleDnaScore::set (SimpleDnaScore.cpp:16-42) over letter classes; letters outside its alphabet
// score 0 (std::map::operator[] default, SimpleDnaScore.cpp:11-14)
int host_score(int x, int y)
{
    if (x == CLS_UNK || y == CLS_UNK) return 0;
    if (x == y) return 3;
    if ((x == CLS_PLUS && y == CLS_GAP) || (x == CLS_GAP && y == CLS_PLUS)) return 3;
    if (x == CLS_PLUS || y == CLS_PLUS) return -4 + -2;
    if (x == CLS_GAP || y == CLS_GAP) return -2;
    return -5;
}

__device__ __forceinline__ int classify(char ch)
{
    switch (ch)
    {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return CLS_GAP;
        case '+': return CLS_PLUS;
    }
    return CLS_UNK;
}

struct MsaArgs
{
    const int* prob;            // [n] problems of this launch (indices into the batch)
    int n;
    const int* prob_seq_off;
    const int* seq_off;
    const char* chars;
    const long long* colchar_off;  // per launch slot
    char* colchar;
    const long long* rows_off;     // per launch slot; row stride = WM
    char* rows;
    int* width;                 // per launch slot
    int* status;                // per launch slot: 0 ok, 1 the profile outgrew the class
    unsigned long long* cells;  // per launch slot
    int WM, LM;                 // columns / letters the class holds
    size_t region;              // bytes of one problem's tables
    unsigned char* gtables;     // tables of the global-memory class (region bytes per slot), else null
};

__host__ __device__ inline size_t msa_region_bytes(int WM, int LM)
{
    size_t b = 0;
    b += sizeof(int) * (size_t)(WM + 1) * (LM + 1);   // SC
    b += sizeof(int) * (size_t)WM * 8 * 2;            // cnt, cs
    b += sizeof(unsigned short) * (size_t)WM * 3;     // ord0, ord1, birth
    b += (size_t)(WM + 1) * (LM + 1);                 // BT
    b += (size_t)WM;                                  // firstgap
    b += (size_t)(WM + LM + 2);                       // ops
    b += (size_t)LM;                                  // seqcls
    b += 16;                                          // W, ncol, err, k
    return (b + 15) & ~(size_t)15;
}

template <int G>
__device__ __forceinline__ int group_scan_incl(int x, int gl, unsigned gmask)
{
#pragma unroll
    for (int o = 1; o < G; o <<= 1)
    {
        const int y = __shfl_up_sync(gmask, x, o, G);
        if (gl >= o) x += y;
    }
    return x;
}

// G lanes per problem, GROUPS problems per CTA; tables in shared memory, or in global memory when a.gtables
template <int G, int GROUPS>
__global__ void __launch_bounds__(G * GROUPS) msa_sp_kernel(MsaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sS[NCLS][8];
    const int tid = threadIdx.x, gl = tid % G, grp = tid / G;
    const int lane = tid & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
    for (int q = tid; q < NCLS * NCLS; q += G * GROUPS) sS[q / NCLS][q % NCLS] = c_S[q / NCLS][q % NCLS];
    __syncthreads();
    const int slot = blockIdx.x * GROUPS + grp;
    if (slot >= a.n) return;  // whole groups leave together; no CTA barrier below
    const int p = a.prob[slot];
    const int WM = a.WM, LM = a.LM;
    unsigned char* base = a.gtables ? a.gtables + (size_t)slot * a.region : smem_raw + (size_t)grp * a.region;
    int* SC = reinterpret_cast<int*>(base);                    // (WM+1) x (LM+1)
    int* cnt = SC + (WM + 1) * (LM + 1);                       // WM x 8   class counts per column id
    int* cs = cnt + WM * 8;                                    // WM x 8   non-gap rows vs letter class
    unsigned short* ord0 = reinterpret_cast<unsigned short*>(cs + WM * 8);
    unsigned short* ord1 = ord0 + WM;
    unsigned short* birth = ord1 + WM;
    unsigned char* BT = reinterpret_cast<unsigned char*>(birth + WM);  // (WM+1) x (LM+1): dir | state<<2
    unsigned char* firstgap = BT + (WM + 1) * (LM + 1);
    unsigned char* ops = firstgap + WM;                        // WM + LM + 2
    unsigned char* seqcls = ops + (WM + LM + 2);               // LM
    int* sv = reinterpret_cast<int*>(base + a.region - 16);    // [0] W, [1] ncol, [2] err, [3] k

    const int s0 = a.prob_seq_off[p], s1 = a.prob_seq_off[p + 1];
    const int nrow = s1 - s0;
    char* colchar = a.colchar + a.colchar_off[slot];
    char* rows = a.rows + a.rows_off[slot];
    unsigned long long cells = 0;

    // ---- first sequence: one column per letter
    {
        const int b = a.seq_off[s0], len = a.seq_off[s0 + 1] - b;
        if (gl == 0) { sv[2] = (len > WM) ? 1 : 0; sv[0] = len; sv[1] = len; }
        __syncwarp(gmask);
        if (!sv[2])
            for (int w = gl; w < len; w += G)
            {
                const char ch = a.chars[b + w];
                const int cl = classify(ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) { cnt[w * 8 + c] = 0; cs[w * 8 + c] = (cl != CLS_GAP && c < NCLS) ? sS[cl][c] : 0; }
                cnt[w * 8 + cl] = 1;
                ord0[w] = (unsigned short)w;
                birth[w] = 0;
                firstgap[w] = (ch == '-');
                colchar[(long long)w * nrow] = ch;
            }
        __syncwarp(gmask);
    }
    unsigned short* ord = ord0;
    unsigned short* ordn = ord1;
    const int S0P = sS[0][CLS_PLUS], S0G = sS[0][CLS_GAP];

    for (int t = 1; t < nrow && !sv[2]; ++t)
    {
        const int b = a.seq_off[s0 + t], len = a.seq_off[s0 + t + 1] - b;
        const int W = sv[0], m = W + 1, n = len + 1, s = t;
        for (int j = gl; j < len; j += G) seqcls[j] = (unsigned char)classify(a.chars[b + j]);
        // ---- borders (MultipleSequenceAlignmentSP.cpp:65-137): the first row in closed form ...
        for (int j = gl; j < n; j += G)
        {
            SC[j] = (j == 0) ? 0 : s * (S0P + (j - 1) * S0G);
            BT[j] = (j == 0) ? (P_MAT | (P_MAT << 2)) : (P_INS | (P_INS << 2));
        }
        // ... the first column as a running sum over the profile, G columns at a time
        {
            int carry = 0;
            for (int i0 = 1; i0 < m; i0 += G)
            {
                const int i = i0 + gl;
                int v = 0;
                if (i < m)
                {
                    const int col = ord[i - 1];
                    v = cs[col * 8 + ((i == 1) ? CLS_PLUS : CLS_GAP)] + cnt[col * 8 + CLS_GAP] * 3;
                }
                const int x = group_scan_incl<G>(v, gl, gmask);
                if (i < m)
                {
                    SC[i * n] = carry + x;
                    BT[i * n] = P_DEL | (P_MAT << 2);  // per-row states here are never "ins" and never asked for "del"
                }
                carry += __shfl_sync(gmask, x, G - 1, G);
            }
        }
        __syncwarp(gmask);
        // ---- anti-diagonal wavefront over the interior (MultipleSequenceAlignmentSP.cpp:139-248)
        for (int d = 2; d <= (m - 1) + (n - 1); ++d)
        {
            const int ilo = max(1, d - (n - 1)), ihi = min(m - 1, d - 1);
            for (int i = ilo + gl; i <= ihi; i += G)
            {
                const int j = d - i;
                const int cj = seqcls[j - 1];
                const int col = ord[i - 1];
                const int ngap = cnt[col * 8 + CLS_GAP];
                const int st_d = BT[(i - 1) * n + (j - 1)] >> 2;
                const int st_l = BT[i * n + (j - 1)] >> 2;
                const int st_u = BT[(i - 1) * n + j] >> 2;
                const int r1 = SC[(i - 1) * n + (j - 1)] + cs[col * 8 + cj] + ngap * sS[st_d == P_INS ? CLS_GAP : CLS_PLUS][cj];
                const int r2 = SC[i * n + (j - 1)] + s * sS[st_l == P_INS ? CLS_GAP : CLS_PLUS][cj];
                const int r3 = SC[(i - 1) * n + j] + cs[col * 8 + (st_u == P_DEL ? CLS_GAP : CLS_PLUS)] + ngap * 3;
                const int fg = firstgap[col];
                int best, bt;
                if (r1 >= r2 && r1 >= r3) { best = r1; bt = P_MAT | ((fg ? P_INS : P_MAT) << 2); }
                else if (r2 >= r1 && r2 >= r3) { best = r2; bt = P_INS | (P_INS << 2); }
                else { best = r3; bt = P_DEL | ((fg ? P_MAT : P_DEL) << 2); }
                SC[i * n + j] = best;
                BT[i * n + j] = (unsigned char)bt;
            }
            __syncwarp(gmask);
        }
        cells += (unsigned long long)(m - 1) * (n - 1);
        // ---- traceback (MultipleSequenceAlignmentSP.cpp:252-301): a serial walk, last operation first
        if (gl == 0)
        {
            int x = m - 1, y = n - 1, k = 0;
            while (x != 0 || y != 0)
            {
                const int dir = BT[x * n + y] & 3;
                ops[k++] = (unsigned char)dir;
                if (dir == P_MAT) { --x; --y; }
                else if (dir == P_INS) --y;
                else --x;
            }
            sv[3] = k;
        }
        __syncwarp(gmask);
        // ---- profile update by all lanes: operation f (in forward order) consumes profile position
        // #(non-insert ops before f) and letter #(non-delete ops before f); an insert opens column ncol + #(inserts before f)
        {
            const int k = sv[3], ncol = sv[1];
            int c_po = 0, c_j = 0, c_ins = 0, err = 0;
            for (int f0 = 0; f0 < k; f0 += G)
            {
                const int f = f0 + gl;
                const int dir = f < k ? ops[k - 1 - f] : -1;
                const int is_ins = dir == P_INS ? 1 : 0, is_del = dir == P_DEL ? 1 : 0, in = f < k ? 1 : 0;
                const int x_po = group_scan_incl<G>(in - is_ins, gl, gmask);
                const int x_j = group_scan_incl<G>(in - is_del, gl, gmask);
                const int x_ins = group_scan_incl<G>(is_ins, gl, gmask);
                if (in)
                {
                    const int po = c_po + x_po - (in - is_ins), j = c_j + x_j - (in - is_del);
                    int col;
                    char ch;
                    if (is_ins)
                    {
                        col = ncol + c_ins + x_ins - 1;
                        if (col >= WM) err = 1;
                        else
                        {
                            ch = a.chars[b + j];
                            const int cl = classify(ch);
#pragma unroll
                            for (int c = 0; c < 8; ++c) { cnt[col * 8 + c] = 0; cs[col * 8 + c] = (cl != CLS_GAP && c < NCLS) ? sS[cl][c] : 0; }
                            cnt[col * 8 + CLS_GAP] = t;
                            cnt[col * 8 + cl] += 1;
                            birth[col] = (unsigned short)t;
                            firstgap[col] = 1;
                        }
                    }
                    else if (!is_del)
                    {
                        col = ord[po];
                        ch = a.chars[b + j];
                        const int cl = classify(ch);
                        cnt[col * 8 + cl] += 1;
                        if (cl != CLS_GAP)
                        {
#pragma unroll
                            for (int c = 0; c < NCLS; ++c) cs[col * 8 + c] += sS[cl][c];
                        }
                    }
                    else
                    {
                        col = ord[po];
                        ch = '-';
                        cnt[col * 8 + CLS_GAP] += 1;
                    }
                    if (!err)
                    {
                        colchar[(long long)col * nrow + t] = ch;
                        ordn[f] = (unsigned short)col;
                    }
                }
                c_po += __shfl_sync(gmask, x_po, G - 1, G);
                c_j += __shfl_sync(gmask, x_j, G - 1, G);
                c_ins += __shfl_sync(gmask, x_ins, G - 1, G);
            }
            err = __any_sync(gmask, err) ? 1 : 0;
            if (gl == 0)
            {
                sv[0] = k;
                sv[1] = ncol + c_ins;
                if (err) sv[2] = 1;
            }
        }
        __syncwarp(gmask);
        unsigned short* tmp = ord; ord = ordn; ordn = tmp;
    }
    __syncwarp(gmask);
    // ---- rows of the final profile (MSA::get, MultipleSequenceAlignment.hpp:56-66)
    if (!sv[2])
    {
        const int W = sv[0];
        for (int k = gl; k < nrow * W; k += G)
        {
            const int t = k / W, w = k % W;
            const int col = ord[w];
            rows[(long long)t * WM + w] = (t >= birth[col]) ? colchar[(long long)col * nrow + t] : '-';
        }
    }
    if (gl == 0)
    {
        a.width[slot] = sv[0];
        a.status[slot] = sv[2];
        a.cells[slot] = cells;
    }
}

void upload_scores()
{
    static bool ready_on[64] = {false};  // per device: __constant__ memory is per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& g_score_ready = ready_on[dev & 63];
    if (g_score_ready) return;
    int S[NCLS][NCLS];
    for (int x = 0; x < NCLS; ++x) for (int y = 0; y < NCLS; ++y) S[x][y] = host_score(x, y);
    RAMBL_CUDA(cudaMemcpyToSymbol(c_S, S, sizeof(S)));
    g_score_ready = true;
}

// size classes: a problem starts in the first class that holds its longest string
// The first class is what real insertion levels look like (homopolymer indels: hundreds of strings of 1-4 letters, a
// profile of 2-4 columns): a 3x3 DP has nothing to spread over lanes, so there a problem is ONE lane -- 128 problems per
// CTA, sorted by their number of strings so that the lanes of a warp run the same number of steps.
struct MsaClass { int LM, WM, G; };
const MsaClass kClasses[] = {{4, 8, 1}, {8, 16, 8}, {16, 32, 8}, {32, 64, 16}, {63, 255, 32}};
constexpr int kNumClasses = 5;  // + the global-memory class, index kNumClasses

template <int G, int GROUPS>
void launch_class(const MsaArgs& a, cudaStream_t st)
{
    const size_t smem = a.gtables ? 16 : a.region * (size_t)GROUPS;
    static size_t configured_on[64] = {0};  // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_on[dev & 63];
    if (smem > configured)
    {
        RAMBL_CUDA(cudaFuncSetAttribute(msa_sp_kernel<G, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    msa_sp_kernel<G, GROUPS><<<(a.n + GROUPS - 1) / GROUPS, G * GROUPS, smem, st>>>(a);
}

}  // namespace

void require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(RAMBL_ERR_CUDA, std::string("rambl_b200 needs a CUDA device (sm_100a); none usable: ") +
                                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
}

void msa_sp_align_batch(const MsaBatch& in, MsaResult& out, cudaStream_t stream)
{
    const int P = (int)in.prob_seq_off.size() - 1;
    out.width.assign(std::max(P, 0), 0);
    out.row_off.assign(std::max(P, 0) + 1, 0);
    out.cap.assign(std::max(P, 0), 0);
    out.rows.clear();
    out.dp_cells = 0;
    out.kernel_ms = 0;
    out.launches = 0;
    if (P <= 0) return;
    require_device();
    upload_scores();

    // ---- per problem: longest string, letters in total -> starting class
    std::vector<int> lmax(P, 1), cls(P, 0);
    std::vector<long long> letters(P, 0);
    std::vector<std::vector<int>> todo(kNumClasses + 1);
    for (int p = 0; p < P; ++p)
    {
        const int s0 = in.prob_seq_off[p], s1 = in.prob_seq_off[p + 1];
        if (s1 <= s0) throw Error(RAMBL_ERR_INVALID, "msa problem without sequences");
        if (s1 - s0 > 65535) throw Error(RAMBL_ERR_CAPACITY, "an insertion level with more than 65535 strings");
        for (int s = s0; s < s1; ++s)
        {
            const int len = in.seq_off[s + 1] - in.seq_off[s];
            lmax[p] = std::max(lmax[p], len);
            letters[p] += len;
        }
        int c = 0;
        while (c < kNumClasses && lmax[p] > kClasses[c].LM) ++c;
        cls[p] = c;
        todo[c].push_back(p);
    }

    const size_t nseq = in.seq_off.size() - 1, nchar = in.chars.size();
    DevBuf<int> d_pso, d_so;
    DevBuf<char> d_chars;
    d_pso.reserve(P + 1); d_so.reserve(nseq + 1); d_chars.reserve(std::max<size_t>(nchar, 1));
    RAMBL_CUDA(cudaMemcpyAsync(d_pso.p, in.prob_seq_off.data(), sizeof(int) * (P + 1), cudaMemcpyHostToDevice, stream));
    RAMBL_CUDA(cudaMemcpyAsync(d_so.p, in.seq_off.data(), sizeof(int) * (nseq + 1), cudaMemcpyHostToDevice, stream));
    if (nchar) RAMBL_CUDA(cudaMemcpyAsync(d_chars.p, in.chars.data(), nchar, cudaMemcpyHostToDevice, stream));

    std::vector<std::vector<char>> prob_rows(P);  // rows of problem p, compact: row t at t * width[p]
    DevBuf<int> d_prob, d_width, d_status;
    DevBuf<long long> d_cco, d_ro;
    DevBuf<unsigned long long> d_cells;
    DevBuf<char> d_colchar, d_rows;
    DevBuf<unsigned char> d_tables;
    cudaEvent_t e0, e1;
    RAMBL_CUDA(cudaEventCreate(&e0));
    RAMBL_CUDA(cudaEventCreate(&e1));
    for (int c = 0; c <= kNumClasses; ++c)
    {
        std::vector<int>& list = todo[c];
        if (list.empty()) continue;
        // longest problems first, neighbours of similar length: the lanes / groups of a warp finish together
        std::sort(list.begin(), list.end(), [&](int x, int y) {
            const int nx = in.prob_seq_off[x + 1] - in.prob_seq_off[x], ny = in.prob_seq_off[y + 1] - in.prob_seq_off[y];
            return nx != ny ? nx > ny : x < y;
        });
        const int n = (int)list.size();
        const bool global = c == kNumClasses;
        int WM, LM, G;
        if (!global) { WM = kClasses[c].WM; LM = kClasses[c].LM; G = kClasses[c].G; }
        else
        {   // tables in global memory: sized for the widest profile any of these problems can reach
            long long wm = 1;
            LM = 1;
            for (int p : list) { wm = std::max(wm, letters[p]); LM = std::max(LM, lmax[p]); }
            if (wm > 65535) throw Error(RAMBL_ERR_CAPACITY, "an insertion level with more than 65535 letters");
            WM = (int)wm;
            G = 32;
        }
        size_t region = msa_region_bytes(WM, LM);
        // one lane per problem: the lanes of a warp walk their regions in step, so the region stride must be an odd
        // number of 4-byte words or all 32 lanes would sit on two shared-memory banks
        if (!global && G == 1) while ((region / 4) % 2 == 0) region += 4;
        std::vector<long long> cco(n + 1, 0), ro(n + 1, 0);
        for (int k = 0; k < n; ++k)
        {
            const long long nrow = in.prob_seq_off[list[k] + 1] - in.prob_seq_off[list[k]];
            cco[k + 1] = cco[k] + (long long)WM * nrow;
            ro[k + 1] = ro[k] + (long long)WM * nrow;
        }
        d_prob.reserve(n); d_width.reserve(n); d_status.reserve(n); d_cells.reserve(n); d_cco.reserve(n + 1); d_ro.reserve(n + 1);
        d_colchar.reserve(std::max<long long>(cco[n], 1));
        d_rows.reserve(std::max<long long>(ro[n], 1));
        if (global) d_tables.reserve(region * (size_t)n);
        RAMBL_CUDA(cudaMemcpyAsync(d_prob.p, list.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream));
        RAMBL_CUDA(cudaMemcpyAsync(d_cco.p, cco.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, stream));
        RAMBL_CUDA(cudaMemcpyAsync(d_ro.p, ro.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, stream));
        MsaArgs a;
        a.prob = d_prob.p; a.n = n; a.prob_seq_off = d_pso.p; a.seq_off = d_so.p; a.chars = d_chars.p;
        a.colchar_off = d_cco.p; a.colchar = d_colchar.p; a.rows_off = d_ro.p; a.rows = d_rows.p;
        a.width = d_width.p; a.status = d_status.p; a.cells = d_cells.p;
        a.WM = WM; a.LM = LM; a.region = region; a.gtables = global ? d_tables.p : nullptr;
        RAMBL_CUDA(cudaEventRecord(e0, stream));
        if (global) launch_class<32, 4>(a, stream);
        else if (G == 1) launch_class<1, 128>(a, stream);
        else if (G == 8) launch_class<8, 16>(a, stream);
        else if (G == 16) launch_class<16, 8>(a, stream);
        else launch_class<32, 2>(a, stream);  // 100 KB of tables per problem: two per SM
        RAMBL_CUDA(cudaEventRecord(e1, stream));
        RAMBL_CUDA(cudaGetLastError());
        out.launches += 1;
        std::vector<int> width(n), status(n);
        std::vector<unsigned long long> cells(n);
        std::vector<char> rows((size_t)ro[n]);
        RAMBL_CUDA(cudaMemcpyAsync(width.data(), d_width.p, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
        RAMBL_CUDA(cudaMemcpyAsync(status.data(), d_status.p, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
        RAMBL_CUDA(cudaMemcpyAsync(cells.data(), d_cells.p, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, stream));
        if (ro[n]) RAMBL_CUDA(cudaMemcpyAsync(rows.data(), d_rows.p, (size_t)ro[n], cudaMemcpyDeviceToHost, stream));
        RAMBL_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        RAMBL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        out.kernel_ms += ms;
        for (int k = 0; k < n; ++k)
        {
            const int p = list[k];
            if (status[k])
            {   // the profile outgrew the class: once more in the next one
                if (global) throw Error(RAMBL_ERR_CAPACITY, "msa profile outgrew the global-memory tables");
                todo[c + 1].push_back(p);
                continue;
            }
            out.dp_cells += cells[k];
            const int nrow = in.prob_seq_off[p + 1] - in.prob_seq_off[p], w = width[k];
            out.width[p] = w;
            out.cap[p] = std::max(w, 1);
            prob_rows[p].resize((size_t)nrow * std::max(w, 1));
            for (int t = 0; t < nrow; ++t)
                memcpy(prob_rows[p].data() + (size_t)t * std::max(w, 1), rows.data() + ro[k] + (long long)t * WM, (size_t)w);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    long long total = 0;
    for (int p = 0; p < P; ++p) { out.row_off[p] = total; total += (long long)prob_rows[p].size(); }
    out.row_off[P] = total;
    out.rows.resize((size_t)total);
    for (int p = 0; p < P; ++p)
        if (!prob_rows[p].empty()) memcpy(&out.rows[(size_t)out.row_off[p]], prob_rows[p].data(), prob_rows[p].size());
}

}  // namespace rambl
