// Device-resident strain walk (kernel in walk.cu): ONE kernel launch runs PartialOrderGraph::streaming_clustering
// (/root/reference/StrainCall/NonparametricClustering.cpp:262-582) for every subgroup of a batch, one CTA per
// subgroup, from the "^" level to the "$" level with no host in the loop:
//
//     per level:  read set + draws  ->  k_inherit copies  ->  log-likelihood update (lines 343-391)
//                 -> weights (50-60, 178-191) -> hard_clustering (17-125) or np_bayes_clustering (127-244)
//                 -> pruning + path extension + the 80-candidate cut (393-551)
//
// The level-synchronous path (engine.cpp + dpm.cu) does the same with the host deciding between levels: one small
// upload, 4-5 launches, one download and a stream synchronisation per level for the whole batch, and every level
// waits for its slowest subgroup.  Here a subgroup's levels are a loop inside its CTA; the level tables are
// uploaded once (the level structure is a property of the graph, not of the candidate strains), subgroups advance
// independently, and the host only closes the result ("$": sort + merge_strains) and runs read_assign.
// A batch with fewer subgroups than SMs gives each subgroup a thread-block CLUSTER: the phases of a level are split over
// all its CTAs and the Gibbs chain runs over distributed shared memory (dpm_dev.cuh).
// Subgroups the kernel cannot take (see WALK_* below) are solved by the level-synchronous path instead.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rambl {

constexpr int WALK_SMAX = 128;    // candidate strains per level the walk kernel handles
constexpr int WALK_KMAX = 1024;   // children of one level before the 80-candidate cut

enum
{
    WALK_DONE = 0,          // reached "$"
    WALK_NO_CANDS = 1,      // every candidate was pruned on the way
    // the level-synchronous path takes over (nothing of the walk is kept):
    WALK_TOO_MANY_STRAINS = 2,
    WALK_TOO_MANY_CHILDREN = 3,
    WALK_TRAIL_FULL = 4,
    WALK_NAN_IN_CUT = 5,
    WALK_OUT_OF_SLOTS = 6,
    WALK_NOT_SETTLED = 7,   // a Gibbs round hit its pass cap (non-finite weights)
    WALK_NOT_RUN = 8
};

struct WalkCand
{
    int slot, node, tail, pad;
    double ab;
    unsigned long long hash, len;
};

struct WalkResult
{
    int status, n_cands, levels, reason_level;
    long long draws, loglik_updates, weight_pairs, gibbs_bytes;
    unsigned long long rounds, passes;
    long long sum_S;                           // candidate strains summed over the Gibbs levels
    int gibbs_levels, unstaged_levels, max_S, pad;
    int free_top, branching, cand_buf, pad2;   // walk state at the stop: free-slot stack height, the branching flag, which candidate buffer
};

struct WalkSub
{
    // ---- the graph (uploaded once)
    const int* label_off;           // [N+1]
    const char* label_chars;
    const int* out_off;             // [N+1]
    const int* out_to;              // [E]
    const int* out_cover;           // [E] reads over the edge (number_of_reads_cover_nodes)
    int end_node;
    int n_levels;                   // levels of the walk; the last one holds "$" alone
    const int* lvl_ent_off;         // [n_levels+1] read-pool entries of a level, in the order the level's nodes list them
    const unsigned char* lvl_dup;   // [n_levels] the most FURTHER entries any read has on the level (0: every read once)
    const unsigned* ent_rid;        // [entries] unique read id; bit 31: a further entry of a read already listed on this level
    const unsigned char* ent_cn;    // [entries] copies
    const char* ent_char1;          // [entries] the entry's letter (levels whose entries are all one letter: nearly all)
    const int* lvl_moff;            // [n_levels] -1, or where the level's entries start in m_soff / m_len (collapsed nodes)
    const unsigned* m_soff;         // offset of an entry's letters in m_chars
    const unsigned char* m_len;     // its length
    const char* m_chars;
    const int* pair_off;            // ReadPairs as CSR over unique reads: one mate id or -1 per copy
    const int* pair_val;
    int R;                          // unique reads == row stride of ll
    int slot_cap;
    // ---- state
    double* ll;                     // [slot_cap][R]
    double* sub;                    // [slot_cap][36]
    unsigned char* present;         // [R] the read has an entry in the strains' log-likelihood maps
    int* free_slots;                // [slot_cap] stack
    WalkCand* cand[2];              // [WALK_KMAX] candidates of this level / the next
    int2* trail;                    // (parent trail index, node)
    int trail_cap;
    // ---- per-level scratch
    double* W;                      // weight tiles + normalisers + letter codes of one level (layout: dpm_dev.cuh)
    int* ent_doff;                  // [max entries of a level + 1] first draw of every entry
    int* draw_entry;                // [max draws of a level]
    int* draw_mate;
    unsigned char* fresh;           // [max entries of a level]
    double* ab_io;                  // [WALK_SMAX] abundances in, increments out
    int2* ops;                      // [WALK_KMAX] slot copies queued by the last extension
    double* kid_ab;                 // [WALK_KMAX] scratch of the cut
    double* lut;                    // [WALK_SMAX][36] log substitution tables of the level's strains
    int* helper;                    // [8] cluster launches: the level descriptor rank 0 publishes (candidates, candidate buffer,
                                    // slot copies, branching flag, stop code or -1)
    // ---- result
    WalkResult* res;
    int* paths;                     // [<= WALK_SMAX][n_levels] node ids of the final candidates' paths
    int* final_slot;                // [WALK_SMAX]
    double* final_ab;               // [WALK_SMAX]
};

struct WalkParams
{
    int n;          // sweep cap (5000)
    int single_buffer;  // bit 0: wide levels may run with one tile buffer (more blocks per round, the tile copy exposed);
                        // bit 1: a level whose weights fit the tile buffers as a whole copies them once and keeps them
    double tau;
    const double* uniforms;               // the std::mt19937(1234) canonical stream
    unsigned long long* counters;         // [0] rounds, [1] passes, [2] rounds that did not settle
};

size_t walk_smem_bytes(int nb, int tile_S, bool cluster);
// launches k_walk<nb> with one CTA -- or one cluster of `cluster` CTAs (2, 4, 8; nb must be 8) -- per subgroup;
// nb = warps per CTA = 32-draw blocks a CTA adds to a Gibbs round (1, 2, 4 or 8)
void launch_walk(const WalkSub* d_subs, int n_subs, const WalkParams& prm, int nb, int tile_S, int cluster, cudaStream_t st,
                 int* launches);

}  // namespace rambl
