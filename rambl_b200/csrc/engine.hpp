// Strain inference over a batch of subgroups (host orchestration of the kernels in walk.cu and dpm.cu).
//
// Replaces, for every subgroup of the batch at once,
//   PartialOrderGraph::infer_strains -> streaming_clustering  (NonparametricClustering.cpp:262-582,704-708)
//   PartialOrderGraph::read_assign (AlignRead form)           (NonparametricClustering.cpp:776-836)
//   the abundance sort of main()                              (StrainCall.cpp:1027)
// Two paths compute the same thing.  The device-resident walk (walk.cu, the default): the level structure of a graph is
// turned into compact tables once, and ONE kernel walks every subgroup from "^" to "$" -- candidate bookkeeping included --
// with one CTA or one thread-block cluster per subgroup; the host closes the results ("$": sort + merge_strains) and
// launches read_assign.  The level-synchronous path (dpm.cu; what the walk cannot take, and the take-over after an early
// "$"): the graph walk and the candidate bookkeeping stay on the host, everything per read x strain runs on the device,
// one set of launches per graph level for all subgroups.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "pog.hpp"

namespace rambl {

struct InferParams
{
    int n = 5000;         // Gibbs sweeps cap, StrainCall.cpp:1021
    float e = 0.01f;      // --error-rate   (held as float by the CLI, StrainCall.cpp:84)
    float tau = 0.02f;    // --tau
    float diff = 0.01f;   // --diff-rate
    bool assign = true;   // run read_assign and the final sort
    bool keep_loglik = false;  // also return the per-read log-likelihood rows of the inferred strains
    int max_cluster = 0;  // > 0: at most this many CTAs per subgroup in the walk (the overlapped solve runs several chunks' kernels side by side)
    bool level_synchronous = false;  // this call: the level-synchronous path only (the side batch of subgroups the walk cannot take)
    // called on the calling thread once the walk kernel is in its stream, i.e. when the host-heavy set-up of the call is
    // over (not called when nothing goes to the walk): the overlapped solve starts the next chunk's host work then
    std::function<void()> on_device_phase;
};

struct StrainResult
{
    double abundance_infer = 0;  // after streaming_clustering (+ merge)
    double abundance = 0;        // after read_assign (== abundance_infer when assign is off)
    std::vector<int> path;       // node ids from "^" to "$"
    double sub[36];              // substitution counts of the strain model
    std::vector<double> loglik;  // [n_reads] when keep_loglik
};

struct SubgroupResult
{
    int status = 0;                     // RAMBL_OK or RAMBL_ERR_NO_STRAINS
    std::vector<StrainResult> strains;  // order of streaming_clustering's result
    std::vector<int> order;             // indices into strains, by abundance (StrainCall.cpp:1027)
    long long draws = 0;                // categorical draws made for this subgroup
    int levels = 0;
};

struct SubgroupInput
{
    const FlatGraph* graph = nullptr;
    std::vector<int> read_cn;             // copy number per unique read
    std::vector<int> pair_off, pair_val;  // ReadPairs as CSR over unique reads (one mate id or -1 per copy)
};

struct EngineStats
{
    int launches = 0;
    int level_steps = 0;
    long long draws = 0;
    long long loglik_updates = 0;  // (read-pool entry, strain) pairs
    float gpu_ms = 0;              // CUDA-event time from the first launch to the last
    float gibbs_ms = 0;            // CUDA-event time spent inside the Gibbs kernel, summed over its launches
    int gibbs_launches = 0;
    long long gibbs_bytes = 0;     // algorithmic bytes of those launches: sweeps x draws x (S weights + 1 uniform) x 8
    long long h2d_bytes = 0, d2h_bytes = 0;
    long long gibbs_rounds = 0, gibbs_passes = 0;  // rounds of 32 speculative draws / passes needed to settle them
    float walk_ms = 0;             // CUDA-event time of the device-resident walk kernel
    int walk_launches = 0;
    int offtable_levels = 0;       // graph levels on which a one-letter strain label can meet a multi-letter read string
    long long walk_bytes = 0;      // its algorithmic bytes: the Gibbs bytes + 16 B per log-likelihood update + 16 B per weight
};

// 1 (default): subgroups are solved by the device-resident walk (walk.cu) where it applies; 0: level-synchronous only
void set_walk_mode(int mode);
int walk_mode();
// warps per CTA of the walk kernel (= 32-draw blocks per Gibbs round): 0 = by batch size, else 1, 2, 4 or 8
void set_walk_blocks(int nb);
int walk_blocks();
// CTAs per subgroup of the walk kernel (a thread-block cluster; the extra CTAs join the Gibbs chains): 0 = by batch
// size, else 1, 2, 4 or 8
void set_walk_cluster(int c);
int walk_cluster();

// What the device-resident walk would do with a subgroup -- a function of the graph and the read pairs alone (the
// level tables are a property of the graph), so it needs no device.
struct WalkEligibility
{
    bool eligible = false;   // the walk takes the subgroup
    bool handoff = false;    // ... up to an early "$" (a level that holds "$" next to other nodes); the level-synchronous path finishes it
    int reason = 0;          // why not: 1 not a DAG, 2 something follows "$" / "$" not reached, 4 entry range (copies or letters > 255),
                             // 5 "^" carries reads, 6 size, 7 mate id out of range, 8 no graph
    int levels = 0;          // graph levels the walk runs (planned so far when not eligible)
    long long entries = 0;   // read-pool entries over all levels (an entry counts once per level its node is listed on)
    int max_entries = 0, max_draws = 0;  // of one level
    int offtable_levels = 0; // levels with a multi-letter read string next to a one-letter node
};
WalkEligibility walk_eligibility(const FlatGraph& g, const SubgroupInput& in);

void infer_batch(const std::vector<SubgroupInput>& in, const InferParams& prm, std::vector<SubgroupResult>& out,
                 EngineStats& stats, cudaStream_t stream = 0);

std::string strain_sequence(const FlatGraph& g, const std::vector<int>& path);        // Strain::strain_seq
std::string strain_plain_sequence(const FlatGraph& g, const std::vector<int>& path);  // Strain::plain_seq

}  // namespace rambl
