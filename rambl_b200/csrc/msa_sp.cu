// Batched progressive sum-of-pairs alignment of insertion strings on sm_100a.
//
// Replaces MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char>::align
// (/root/reference/StrainCall/MultipleSequenceAlignment.hpp:87-107,
//  MultipleSequenceAlignmentSP.cpp:10-301) as PartialOrderGraph::canonize_insert_at_level calls it
// (PartialOrderGraph.cpp:449-455,507-519,546): every graph level whose insertions differ in length is
// one PROBLEM; all problems of all subgroups go through one launch, one warp per problem.
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
// Layout: profile columns have a stable id (creation order); `ord` maps profile position -> id, so a
// column insertion only shifts `ord`.  Letters live column-major in global scratch (colchar[id][row]);
// rows above a column's birth row are '-' implicitly.  DP tables, class counts and `ord` sit in
// shared memory; the wavefront runs one anti-diagonal per step across the lanes of the warp.
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
    const int* prob_seq_off;
    const int* seq_off;
    const char* chars;
    const long long* colchar_off;
    char* colchar;
    const long long* rows_off;
    char* rows;
    const int* cap;
    int* width;
    int* status;
    unsigned long long* cells;
    int smem_W;  // largest cap in the batch
    int smem_L;  // longest sequence in the batch
};

__global__ void __launch_bounds__(32) msa_sp_kernel(MsaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int p = blockIdx.x, lane = threadIdx.x;
    const int WM = a.smem_W, LM = a.smem_L;
    int* SC = reinterpret_cast<int*>(smem_raw);                // (WM+1) x (LM+1)
    int* cnt = SC + (WM + 1) * (LM + 1);                       // WM x 8   (class counts per column id)
    int* cs = cnt + WM * 8;                                    // WM x 8   (non-gap rows vs letter class)
    unsigned short* ord0 = reinterpret_cast<unsigned short*>(cs + WM * 8);
    unsigned short* ord1 = ord0 + WM;
    unsigned short* birth = ord1 + WM;
    unsigned char* BT = reinterpret_cast<unsigned char*>(birth + WM);  // (WM+1) x (LM+1): dir | state<<2
    unsigned char* firstgap = BT + (WM + 1) * (LM + 1);
    unsigned char* ops = firstgap + WM;                        // WM + LM + 2
    unsigned char* seqcls = ops + (WM + LM + 2);               // LM
    __shared__ int sh_W, sh_ncol, sh_err;

    const int s0 = a.prob_seq_off[p], s1 = a.prob_seq_off[p + 1];
    const int nrow = s1 - s0;
    const int cap = a.cap[p];
    char* colchar = a.colchar + a.colchar_off[p];
    char* rows = a.rows + a.rows_off[p];
    unsigned long long cells = 0;

    // ---- first sequence: one column per letter
    {
        const int b = a.seq_off[s0], len = a.seq_off[s0 + 1] - b;
        if (lane == 0) { sh_err = (len > cap) ? 1 : 0; sh_W = len; sh_ncol = len; }
        __syncwarp();
        if (!sh_err)
            for (int w = lane; w < len; w += 32)
            {
                const char ch = a.chars[b + w];
                for (int c = 0; c < 8; ++c) cnt[w * 8 + c] = 0;
                cnt[w * 8 + classify(ch)] = 1;
                ord0[w] = (unsigned short)w;
                birth[w] = 0;
                firstgap[w] = (ch == '-');
                colchar[(long long)w * nrow] = ch;
            }
        __syncwarp();
    }
    unsigned short* ord = ord0;
    unsigned short* ordn = ord1;

    for (int t = 1; t < nrow && !sh_err; ++t)
    {
        const int b = a.seq_off[s0 + t], len = a.seq_off[s0 + t + 1] - b;
        const int W = sh_W, m = W + 1, n = len + 1, s = t;
        for (int j = lane; j < len; j += 32) seqcls[j] = (unsigned char)classify(a.chars[b + j]);
        // per-column sums of the non-gap rows against every letter class
        for (int k = lane; k < W * NCLS; k += 32)
        {
            const int col = ord[k / NCLS], y = k % NCLS;
            int v = 0;
#pragma unroll
            for (int c = 0; c < NCLS; ++c)
                if (c != CLS_GAP) v += cnt[col * 8 + c] * c_S[c][y];
            cs[col * 8 + y] = v;
        }
        __syncwarp();
        // borders (MultipleSequenceAlignmentSP.cpp:65-137)
        if (lane == 0)
        {
            SC[0] = 0;
            BT[0] = P_MAT | (P_MAT << 2);
            for (int j = 1; j < n; ++j)
            {
                SC[j] = SC[j - 1] + s * c_S[0][j == 1 ? CLS_PLUS : CLS_GAP];
                BT[j] = P_INS | (P_INS << 2);
            }
        }
        else if (lane == 1)
        {
            int acc = 0;
            for (int i = 1; i < m; ++i)
            {
                const int col = ord[i - 1], y = (i == 1) ? CLS_PLUS : CLS_GAP;
                acc += cs[col * 8 + y] + cnt[col * 8 + CLS_GAP] * 3;
                SC[i * n] = acc;
                BT[i * n] = P_DEL | (P_MAT << 2);  // per-row states here are never "ins" and never asked for "del"
            }
        }
        __syncwarp();
        // anti-diagonal wavefront over the interior (MultipleSequenceAlignmentSP.cpp:139-248)
        for (int d = 2; d <= (m - 1) + (n - 1); ++d)
        {
            const int ilo = max(1, d - (n - 1)), ihi = min(m - 1, d - 1);
            for (int i = ilo + lane; i <= ihi; i += 32)
            {
                const int j = d - i;
                const int cj = seqcls[j - 1];
                const int col = ord[i - 1];
                const int ngap = cnt[col * 8 + CLS_GAP];
                const int st_d = BT[(i - 1) * n + (j - 1)] >> 2;
                const int st_l = BT[i * n + (j - 1)] >> 2;
                const int st_u = BT[(i - 1) * n + j] >> 2;
                const int r1 = SC[(i - 1) * n + (j - 1)] + cs[col * 8 + cj] + ngap * c_S[st_d == P_INS ? CLS_GAP : CLS_PLUS][cj];
                const int r2 = SC[i * n + (j - 1)] + s * c_S[st_l == P_INS ? CLS_GAP : CLS_PLUS][cj];
                const int r3 = SC[(i - 1) * n + j] + cs[col * 8 + (st_u == P_DEL ? CLS_GAP : CLS_PLUS)] + ngap * 3;
                const int fg = firstgap[col];
                int best, bt;
                if (r1 >= r2 && r1 >= r3) { best = r1; bt = P_MAT | ((fg ? P_INS : P_MAT) << 2); }
                else if (r2 >= r1 && r2 >= r3) { best = r2; bt = P_INS | (P_INS << 2); }
                else { best = r3; bt = P_DEL | ((fg ? P_MAT : P_DEL) << 2); }
                SC[i * n + j] = best;
                BT[i * n + j] = (unsigned char)bt;
            }
            __syncwarp();
        }
        cells += (unsigned long long)(m - 1) * (n - 1);
        // traceback and profile update (MultipleSequenceAlignmentSP.cpp:252-301)
        if (lane == 0)
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
            int po = 0, pn = 0, j = 0, ncol = sh_ncol, err = 0;
            for (int q = k - 1; q >= 0; --q)
            {
                const int dir = ops[q];
                int col;
                char ch;
                if (dir == P_INS)
                {
                    if (ncol >= cap) { err = 1; break; }
                    col = ncol++;
                    ch = a.chars[b + j++];
                    for (int c = 0; c < 8; ++c) cnt[col * 8 + c] = 0;
                    cnt[col * 8 + CLS_GAP] = t;
                    cnt[col * 8 + classify(ch)] += 1;
                    birth[col] = (unsigned short)t;
                    firstgap[col] = 1;
                }
                else if (dir == P_MAT)
                {
                    col = ord[po++];
                    ch = a.chars[b + j++];
                    cnt[col * 8 + classify(ch)] += 1;
                }
                else
                {
                    col = ord[po++];
                    ch = '-';
                    cnt[col * 8 + CLS_GAP] += 1;
                }
                colchar[(long long)col * nrow + t] = ch;
                ordn[pn++] = (unsigned short)col;
            }
            sh_W = pn;
            sh_ncol = ncol;
            if (err) sh_err = 1;
        }
        __syncwarp();
        unsigned short* tmp = ord; ord = ordn; ordn = tmp;
    }
    __syncwarp();
    // ---- rows of the final profile (MSA::get, MultipleSequenceAlignment.hpp:56-66)
    if (!sh_err)
    {
        const int W = sh_W;
        for (int k = lane; k < nrow * W; k += 32)
        {
            const int t = k / W, w = k % W;
            const int col = ord[w];
            rows[(long long)t * cap + w] = (t >= birth[col]) ? colchar[(long long)col * nrow + t] : '-';
        }
    }
    if (lane == 0)
    {
        a.width[p] = sh_W;
        a.status[p] = sh_err;
        a.cells[p] = cells;
    }
}

size_t msa_smem_bytes(int WM, int LM)
{
    size_t b = 0;
    b += sizeof(int) * (size_t)(WM + 1) * (LM + 1);
    b += sizeof(int) * (size_t)WM * 8 * 2;
    b += sizeof(unsigned short) * (size_t)WM * 3;
    b += (size_t)(WM + 1) * (LM + 1);
    b += (size_t)WM;
    b += (size_t)(WM + LM + 2);
    b += (size_t)LM;
    return (b + 15) & ~(size_t)15;
}

bool g_score_ready = false;
void upload_scores()
{
    if (g_score_ready) return;
    int S[NCLS][NCLS];
    for (int x = 0; x < NCLS; ++x) for (int y = 0; y < NCLS; ++y) S[x][y] = host_score(x, y);
    RAMBL_CUDA(cudaMemcpyToSymbol(c_S, S, sizeof(S)));
    g_score_ready = true;
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
    if (P <= 0) return;
    require_device();
    upload_scores();

    std::vector<long long> colchar_off(P + 1, 0), rows_off(P + 1, 0);
    int WM = 1, LM = 1;
    for (int p = 0; p < P; ++p)
    {
        const int s0 = in.prob_seq_off[p], s1 = in.prob_seq_off[p + 1];
        if (s1 <= s0) throw Error(RAMBL_ERR_INVALID, "msa problem without sequences");
        long long total = 0;
        for (int s = s0; s < s1; ++s)
        {
            const int len = in.seq_off[s + 1] - in.seq_off[s];
            if (len > MSA_LMAX) throw Error(RAMBL_ERR_CAPACITY, "insertion longer than MSA_LMAX letters");
            LM = std::max(LM, len);
            total += len;
        }
        const int cap = (int)std::min<long long>(std::max<long long>(total, 1), MSA_WMAX);
        out.cap[p] = cap;
        WM = std::max(WM, cap);
        const long long nrow = s1 - s0;
        colchar_off[p + 1] = colchar_off[p] + (long long)cap * nrow;
        rows_off[p + 1] = rows_off[p] + (long long)cap * nrow;
    }
    for (int p = 0; p <= P; ++p) out.row_off[p] = rows_off[p];
    const size_t smem = msa_smem_bytes(WM, LM);
    if (smem > 200 * 1024) throw Error(RAMBL_ERR_CAPACITY, "msa problem does not fit shared memory");

    const size_t nseq = in.seq_off.size() - 1, nchar = in.chars.size();
    DevBuf<int> d_pso, d_so, d_cap, d_width, d_status;
    DevBuf<char> d_chars, d_colchar, d_rows;
    DevBuf<long long> d_cco, d_ro;
    DevBuf<unsigned long long> d_cells;
    d_pso.reserve(P + 1); d_so.reserve(nseq + 1); d_cap.reserve(P); d_width.reserve(P); d_status.reserve(P);
    d_chars.reserve(std::max<size_t>(nchar, 1)); d_colchar.reserve(std::max<long long>(colchar_off[P], 1));
    d_rows.reserve(std::max<long long>(rows_off[P], 1)); d_cco.reserve(P + 1); d_ro.reserve(P + 1); d_cells.reserve(P);
    RAMBL_CUDA(cudaMemcpyAsync(d_pso.p, in.prob_seq_off.data(), sizeof(int) * (P + 1), cudaMemcpyHostToDevice, stream));
    RAMBL_CUDA(cudaMemcpyAsync(d_so.p, in.seq_off.data(), sizeof(int) * (nseq + 1), cudaMemcpyHostToDevice, stream));
    if (nchar) RAMBL_CUDA(cudaMemcpyAsync(d_chars.p, in.chars.data(), nchar, cudaMemcpyHostToDevice, stream));
    RAMBL_CUDA(cudaMemcpyAsync(d_cap.p, out.cap.data(), sizeof(int) * P, cudaMemcpyHostToDevice, stream));
    RAMBL_CUDA(cudaMemcpyAsync(d_cco.p, colchar_off.data(), sizeof(long long) * (P + 1), cudaMemcpyHostToDevice, stream));
    RAMBL_CUDA(cudaMemcpyAsync(d_ro.p, rows_off.data(), sizeof(long long) * (P + 1), cudaMemcpyHostToDevice, stream));

    MsaArgs a;
    a.prob_seq_off = d_pso.p; a.seq_off = d_so.p; a.chars = d_chars.p;
    a.colchar_off = d_cco.p; a.colchar = d_colchar.p; a.rows_off = d_ro.p; a.rows = d_rows.p;
    a.cap = d_cap.p; a.width = d_width.p; a.status = d_status.p; a.cells = d_cells.p;
    a.smem_W = WM; a.smem_L = LM;
    RAMBL_CUDA(cudaFuncSetAttribute(msa_sp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    RAMBL_CUDA(cudaEventCreate(&e0));
    RAMBL_CUDA(cudaEventCreate(&e1));
    RAMBL_CUDA(cudaEventRecord(e0, stream));
    msa_sp_kernel<<<P, 32, smem, stream>>>(a);
    RAMBL_CUDA(cudaEventRecord(e1, stream));
    RAMBL_CUDA(cudaGetLastError());

    out.rows.resize((size_t)rows_off[P]);
    std::vector<int> status(P);
    std::vector<unsigned long long> cells(P);
    RAMBL_CUDA(cudaMemcpyAsync(out.width.data(), d_width.p, sizeof(int) * P, cudaMemcpyDeviceToHost, stream));
    RAMBL_CUDA(cudaMemcpyAsync(status.data(), d_status.p, sizeof(int) * P, cudaMemcpyDeviceToHost, stream));
    RAMBL_CUDA(cudaMemcpyAsync(cells.data(), d_cells.p, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, stream));
    if (rows_off[P]) RAMBL_CUDA(cudaMemcpyAsync(&out.rows[0], d_rows.p, (size_t)rows_off[P], cudaMemcpyDeviceToHost, stream));
    RAMBL_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    RAMBL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out.kernel_ms = ms;
    out.launches = 1;
    for (int p = 0; p < P; ++p)
    {
        if (status[p]) throw Error(RAMBL_ERR_CAPACITY, "msa profile grew past MSA_WMAX columns");
        out.dp_cells += cells[p];
    }
}

}  // namespace rambl
