// Device side of the Dirichlet-process strain clustering (kernels in dpm.cu).
//
// One LEVEL STEP processes, for every subgroup of the batch at once, what the reference does at
// the end of a graph level inside PartialOrderGraph::streaming_clustering
// (/root/reference/StrainCall/NonparametricClustering.cpp:336-458):
//   k_loglik   per-read log-likelihood update of every candidate strain   (lines 343-391)
//   k_weights  exp(loglik(read) + loglik(mate)) per (draw, strain)        (lines 50-60, 178-191)
//   k_hard     soft assignment + sufficient statistics + model update     (hard_clustering, 17-125)
//   k_gibbs_w  sequential Gibbs sweeps + statistics + model update        (np_bayes_clustering, 127-244)
//   k_gibbs    the same for levels of more than 128 strains (and wide levels of big batches)
//   k_inherit  child strains take a copy of their parent's state          (Strain copies, 505-522)
// A "group" is one subgroup's share of the step; all groups of a step sit in one descriptor array.
#pragma once
#include <cuda_runtime.h>

namespace rambl {

constexpr int DPM_SMAX = 256;  // candidate strains per subgroup and level (the reference prunes to ~80)

enum { MODE_NONE = 0, MODE_HARD = 1, MODE_GIBBS = 2, MODE_ASSIGN = 3 };

struct StepGroup
{
    double* ll;              // [slot][read] log-likelihood matrix of the subgroup
    long long ll_stride;     // reads per row
    double* sub;             // [slot][36] substitution counts (rows A,C,G,T,-,= ; columns the same)
    const char* label_chars; // node labels of the subgroup's graph
    const char* pool_chars;  // read-pool strings of the subgroup's graph
    int S;                   // candidate strains
    int m;                   // read-pool entries at this level
    int D;                   // draws = sum of copy numbers
    int mode;
    int nsweeps;
    int read_size;
    // offsets into the step's int arena
    int slot_off;            // [S] slot of strain s
    int lab_off;             // [S] label offset, [S] label length (consecutive)
    int rid_off;             // [m] read id, then [m] string offset, [m] string length, [m] is-new flag
    int draw_off;            // [D] index of the level read, then [D] effective mate id or -1
    // offsets into the step's double arena
    int ab_off;              // in: [S] abundance of strain s;  out: [S] abundance increment
    long long w_off;         // into the weights scratch, [D][S]
};

struct InheritOp
{
    double* ll;
    long long ll_stride;
    double* sub;
    int src, dst;
};

struct StepLaunch
{
    const StepGroup* groups;  // device
    int n_groups;
    const int* iarena;        // device
    double* darena;           // device (in/out)
    double* weights;          // device scratch
    const double* uniforms;   // device, the std::mt19937(1234) canonical stream
    int n_uniforms;
    int max_S, max_m, max_D;
    bool any_hard, any_gibbs;
    unsigned long long* counters = nullptr;  // device, [0] rounds of 32 draws, [1] passes over those rounds
    cudaEvent_t gibbs_begin = nullptr, gibbs_end = nullptr;  // optional: bracket the k_gibbs launch
};

void set_gibbs_blocks(int blocks);  // 0 = automatic, else 1, 2 or 4 blocks of 32 draws per round
void launch_level_step(const StepLaunch& L, cudaStream_t st, int* launches);
void launch_inherit(const InheritOp* d_ops, int n_ops, long long max_stride, cudaStream_t st, int* launches);
void launch_init_models(double* sub, int n_slots, double e, cudaStream_t st, int* launches);

}  // namespace rambl
