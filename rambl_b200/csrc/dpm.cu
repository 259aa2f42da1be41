// Kernels of the Dirichlet-process strain clustering for sm_100a (see dpm.cuh for the mapping to
// /root/reference/StrainCall/NonparametricClustering.cpp and Strain.cpp).
//
// Numerics: the reference works in x87 long double; the device works in FP64.  A read's weight
// under a strain is kept as exp(loglik) and multiplied by the strain's current mass, where the
// reference forms exp(log(mass/total) + loglik); both feed std::discrete_distribution, which
// normalises its weights, so the two differ by rounding only (DESIGN.md states the tolerance
// and the parity tests check paths, assignments and abundances against the oracle).
#include "dpm.cuh"

#include <cmath>

#include "common.hpp"

namespace rambl {

namespace {

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
        atomicAdd(&row[rid], d);  // Strain::update_read_loglik; rows start at 0, so "create" == "add"
    }
}

__global__ void __launch_bounds__(128) k_weights(const StepGroup* __restrict__ groups, const int* __restrict__ I,
                                                 double* __restrict__ W)
{
    const StepGroup g = groups[blockIdx.y];
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.mode == MODE_NONE || d >= g.D) return;
    const int r = I[g.draw_off + d];
    const int mate = I[g.draw_off + g.D + d];
    const int rid = I[g.rid_off + r];
    double* w = W + g.w_off + (long long)d * g.S;
    for (int s = 0; s < g.S; ++s)
    {
        const double* row = g.ll + (long long)I[g.slot_off + s] * g.ll_stride;
        double v = row[rid];
        if (mate >= 0) v += row[mate];
        w[s] = exp(v);
    }
}

// hard_clustering, NonparametricClustering.cpp:17-125: every read copy is spread over the strains by
// its posterior; masses and substitution counts are summed and folded into the strain models.
__global__ void __launch_bounds__(256) k_hard(const StepGroup* __restrict__ groups, const int* __restrict__ I,
                                              double* __restrict__ Dar, const double* __restrict__ W)
{
    const StepGroup g = groups[blockIdx.x];
    if (g.mode != MODE_HARD) return;
    __shared__ double acc[DPM_SMAX * 37];
    __shared__ double ab[DPM_SMAX];
    const int tid = threadIdx.x, S = g.S;
    for (int k = tid; k < S * 37; k += blockDim.x) acc[k] = 0;
    for (int s = tid; s < S; s += blockDim.x) ab[s] = Dar[g.ab_off + s];
    __syncthreads();
    for (int d = tid; d < g.D; d += blockDim.x)
    {
        const double* w = W + g.w_off + (long long)d * S;
        double T = 0;
        for (int s = 0; s < S; ++s) T += ab[s] * w[s];
        const int r = I[g.draw_off + d];
        const char* rs = g.pool_chars + I[g.rid_off + g.m + r];
        const int rl = I[g.rid_off + 2 * g.m + r];
        const int isnew = I[g.rid_off + 3 * g.m + r];
        for (int s = 0; s < S; ++s)
        {
            const double p = ab[s] * w[s] / T;
            atomicAdd(&acc[s * 37], p);
            const char* lab = g.label_chars + I[g.lab_off + s];
            const int lab_l = I[g.lab_off + S + s];
            if (rl == 1)
            {
                if (lab_l == 1)
                {
                    const int a = letter_code(lab[0]), b = letter_code(rs[0]);
                    if (a < 6 && b < 6) atomicAdd(&acc[s * 37 + 1 + a * 6 + b], p);
                }
            }
            else if (isnew)
            {
                int ii = lab_l, jj = rl;
                while (ii > 0 && jj > 0)
                {
                    const int a = letter_code(lab[--ii]), b = letter_code(rs[--jj]);
                    if (a < 6 && b < 6) atomicAdd(&acc[s * 37 + 1 + a * 6 + b], p);
                }
            }
            else
            {
                int ii = 0, jj = 0;
                while (ii < lab_l && jj < rl)
                {
                    const int a = letter_code(lab[ii++]), b = letter_code(rs[jj++]);
                    if (a < 6 && b < 6) atomicAdd(&acc[s * 37 + 1 + a * 6 + b], p);
                }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < S * 36; k += blockDim.x)
    {
        const int s = k / 36, q = k % 36;
        const double v = acc[s * 37 + 1 + q];
        if (v != 0) g.sub[(long long)I[g.slot_off + s] * 36 + q] += v;
    }
    for (int s = tid; s < S; s += blockDim.x) Dar[g.ab_off + s] = acc[s * 37];
}

// np_bayes_clustering (NonparametricClustering.cpp:127-244) and read_assign (776-836): a sequential
// Gibbs chain.  One warp per subgroup; lane l owns strains l, l+32, l+64, l+96.  Each draw multiplies
// the strain masses by the read's weights, takes an inclusive scan in strain order, and picks the
// first strain whose cumulative weight reaches u * total -- std::discrete_distribution's
// lower_bound over the normalised partial sums (bits/random.tcc), with its rule that fewer than
// two weights consume no random number.  u comes from the shared std::mt19937(1234) stream, which
// the reference restarts on every call.
__global__ void __launch_bounds__(32) k_gibbs(const StepGroup* __restrict__ groups, const int* __restrict__ I,
                                              double* __restrict__ Dar, const double* __restrict__ W,
                                              const double* __restrict__ U)
{
    const StepGroup g = groups[blockIdx.x];
    if (g.mode != MODE_GIBBS && g.mode != MODE_ASSIGN) return;
    __shared__ int cnt[DPM_SMAX * 8];
    const int lane = threadIdx.x, S = g.S;
    const unsigned full = 0xffffffffu;
    for (int k = lane; k < S * 8; k += 32) cnt[k] = 0;
    double a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = (lane + 32 * k < S) ? Dar[g.ab_off + lane + 32 * k] : 0.0;
    __syncwarp();
    const int nblk = (S + 31) / 32;
    int ui = 0;
    for (int sweep = 0; sweep < g.nsweeps; ++sweep)
    {
        for (int d = 0; d < g.D; ++d)
        {
            const double* w = W + g.w_off + (long long)d * S;
            double cum[4];
            double carry = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                if (k < nblk)
                {
                    const int s = lane + 32 * k;
                    double x = (s < S) ? a[k] * w[s] : 0.0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const double y = __shfl_up_sync(full, x, o);
                        if (lane >= o) x += y;
                    }
                    cum[k] = x + carry;
                    carry = __shfl_sync(full, cum[k], 31);
                }
            }
            int c = 0;
            if (S >= 2)
            {
                const double t = U[ui++] * carry;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nblk)
                    {
                        const bool below = (lane + 32 * k < S) && (cum[k] < t);
                        c += __popc(__ballot_sync(full, below));
                    }
                if (c > S - 1) c = S - 1;
            }
            if (lane == (c & 31))
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k == (c >> 5)) a[k] += 1.0;
                if (g.mode == MODE_GIBBS)
                {
                    const int r = I[g.draw_off + d];
                    const int rl = I[g.rid_off + 2 * g.m + r];
                    const int b = (rl == 1) ? letter_code(g.pool_chars[I[g.rid_off + g.m + r]]) : 7;
                    cnt[c * 8 + b] += 1;
                }
            }
        }
    }
    // normalise the masses; fold the averaged counts into the models (lines 217-243)
    double z = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) z += a[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(full, z, o);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const int s = lane + 32 * k;
        if (s >= S) continue;
        double v = a[k] / z;
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
        k_weights<<<dim3((L.max_D + 127) / 128, L.n_groups), 128, 0, st>>>(L.groups, L.iarena, L.weights);
        ++*launches;
    }
    if (L.any_hard)
    {
        k_hard<<<L.n_groups, 256, 0, st>>>(L.groups, L.iarena, L.darena, L.weights);
        ++*launches;
    }
    if (L.any_gibbs)
    {
        k_gibbs<<<L.n_groups, 32, 0, st>>>(L.groups, L.iarena, L.darena, L.weights, L.uniforms);
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
