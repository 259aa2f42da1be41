/* rambl_b200.h -- C ABI of the B200-native StrainCall hot path.
 *
 * Drop-in boundary for RAMBL's StrainCall (reference: homopolymer/RAMBL, StrainCall/).  Every entry
 * point names the reference interface it replaces.  Plain pointers and sizes only; all functions
 * return 0 (RAMBL_OK) or a positive error code, and rambl_last_error() gives the message.  There is
 * no CPU fallback: anything that needs the device fails with RAMBL_ERR_CUDA when none is usable.
 *
 * A "subgroup" is one StrainCall problem: a gene window plus the reads mapped to it, in the form
 * StrainCall holds them after load_mapping_reads (StrainCall.cpp:480-670): de-duplicated AlignRead
 * tuples <relative_pos, cigar, seq, "", copies> and ReadPairs uid -> one mate uid (or -1) per copy.
 * A "batch" is any number of subgroups that are built and solved together (scripts/rambl.py runs
 * one StrainCall process per seed gene, rambl.py:165-194; here they share the device launches).
 */
#ifndef RAMBL_B200_H
#define RAMBL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    RAMBL_OK = 0,
    RAMBL_ERR_CUDA = 1,       /* no usable sm_100 device, or a CUDA call failed */
    RAMBL_ERR_INVALID = 2,    /* malformed input */
    RAMBL_ERR_CAPACITY = 3,   /* a problem exceeds a compiled-in kernel limit */
    RAMBL_ERR_NO_STRAINS = 4, /* every candidate strain was pruned (the reference is undefined here) */
    RAMBL_ERR_STATE = 5       /* calls made in the wrong order */
};

typedef struct rambl_batch rambl_batch;

typedef struct rambl_stats
{
    int32_t gpu_launches;     /* kernels launched by this library since the batch was created */
    int32_t level_steps;      /* level-synchronous steps of the strain search */
    int64_t draws;            /* categorical draws of the Gibbs sweeps */
    int64_t loglik_updates;   /* (read-pool entry, strain) log-likelihood updates */
    int64_t msa_dp_cells;     /* (profile column x letter) cells of the insertion alignments */
    int32_t msa_problems;
    float msa_kernel_ms;      /* CUDA-event time of the alignment kernel */
    float infer_gpu_ms;       /* CUDA-event time of rambl_batch_infer, first launch to last */
    int64_t h2d_bytes;        /* bytes copied host->device / device->host by the library */
    int64_t d2h_bytes;
    float gibbs_kernel_ms;    /* CUDA-event time inside the Gibbs-sweep kernel, summed over its launches */
    int32_t gibbs_launches;
    int64_t gibbs_alg_bytes;  /* sweeps x draws x (S weights + 1 uniform) x 8 bytes over those launches */
    int64_t gibbs_rounds;     /* rounds of 32 speculative draws, and the passes it took to settle them */
    int64_t gibbs_passes;
    float dpm_kernel_ms;      /* CUDA-event time of the device-resident strain walk (one kernel, one CTA per subgroup) */
    int32_t dpm_launches;
    int64_t dpm_alg_bytes;    /* its algorithmic bytes: Gibbs bytes + 16 B per log-likelihood update + 16 B per weight */
    int32_t offtable_levels;  /* graph levels where a one-letter strain label can meet a multi-letter read string: the reference
                               * then counts substitution keys outside its 6x6 table; this library does not (0 on every test input) */
    int32_t reserved;
} rambl_stats;

const char* rambl_last_error(void);
int rambl_device_count(void);
/* Select the CUDA device of the calling thread (cudaSetDevice); one process per GPU is the intended layout. */
int rambl_set_device(int32_t device);
void rambl_free(void* p); /* for every char* this library returns */
/* Device and pinned-host buffers are cached between calls (cudaMalloc/cudaFree cost up to a second per
 * strain search), and so are the large pageable host arrays of the flat graphs (fresh pages were a fifth of graph
 * construction; at most $RAMBL_HOST_CACHE_MB, default 16384, the blocks released longest ago go first); this returns all
 * cached blocks to the driver / the allocator. */
void rambl_release_cached_memory(void);
/* Bytes of pageable host memory the cache holds right now (released by batches, not yet reused). */
int64_t rambl_cached_host_bytes(void);
/* The Gibbs-sweep kernels take 1, 2, 4 or 8 blocks of 32 draws per round; 0 (the default) lets the library
 * choose by batch size and strain count.  -1, -2, -4 pin the block count AND the four-warps-per-block kernel
 * that otherwise serves only levels of more than 64 strains.  Every setting computes the same chain -- this
 * is a measurement and test hook, not a results knob.  Returns RAMBL_ERR_INVALID for any other value. */
int rambl_set_gibbs_blocks(int32_t blocks);
/* The strain search of rambl_batch_infer runs as ONE kernel that walks every subgroup's graph levels on the device
 * (mode 1, the default).  Mode 0 keeps the level-synchronous path only (host decisions between levels), which also
 * solves the subgroups the walk kernel cannot take.  rambl_set_walk_blocks pins the warps per subgroup of the walk
 * kernel (1, 2, 4, 8; 0 = by batch size).  Both paths compute the same chains: measurement and test hooks. */
int rambl_set_walk_mode(int32_t mode);
int rambl_set_walk_blocks(int32_t blocks);
/* CTAs per subgroup of the walk kernel: a thread-block cluster whose extra CTAs join the Gibbs chains over
 * distributed shared memory (1, 2, 4, 8; 0 = by batch size: clusters when the batch has fewer subgroups than SMs). */
int rambl_set_walk_cluster(int32_t ctas);
/* Host worker threads for the per-subgroup host work (graph construction, staging).  0 = $RAMBL_HOST_THREADS if
 * set, else one per hardware thread.  scripts/rambl.py's `--cores` (rambl.py:179) is the natural value; several
 * processes sharing a box (one per GPU) should each take their share. */
int rambl_set_host_threads(int32_t n);

/* ---- MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char>::align
 *      (MultipleSequenceAlignment.hpp:87-107, MultipleSequenceAlignmentSP.cpp:10-301), batched.
 * Problem p aligns sequences prob_seq_off[p] .. prob_seq_off[p+1]-1 in that order; sequence s is
 * letters[seq_off[s] .. seq_off[s+1]).  On return width[p] is the profile width (MSA::size()) and row
 * t of problem p (MSA::get(t)) is rows[row_off[p] + t*row_stride[p] .. + width[p]).  rows must hold
 * rambl_msa_rows_capacity() bytes. dp_cells / kernel_ms may be NULL. */
int64_t rambl_msa_rows_capacity(int32_t n_problems, const int32_t* prob_seq_off, const int32_t* seq_off);
int rambl_msa_sp_align_batch(int32_t n_problems, const int32_t* prob_seq_off, const int32_t* seq_off,
                             const char* letters, int32_t* width, int64_t* row_off, int32_t* row_stride, char* rows,
                             uint64_t* dp_cells, float* kernel_ms);

/* ---- batches */
rambl_batch* rambl_batch_create(void);
void rambl_batch_destroy(rambl_batch* b);

/* One subgroup = the arguments of PartialOrderGraph(GenomeSeq& G, vector<AlignRead>& R)
 * (PartialOrderGraph.hpp:236) plus the ReadPairs later given to infer_strains / read_assign
 * (PartialOrderGraph.hpp:335-340) as CSR: mates of unique read u are pair_val[pair_off[u] ..
 * pair_off[u+1]), one per copy.  pair_off/pair_val may be NULL (no read is paired).
 * Returns the subgroup index (>= 0) or -error. */
int rambl_batch_add_subgroup(rambl_batch* b, const char* gene, int32_t n_reads, const int32_t* pos,
                             const char* const* cigar, const char* const* seq, const int32_t* copies,
                             const int32_t* pair_off, const int32_t* pair_val);

/* The same with the CIGAR and read strings packed: read i has CIGAR cigar_chars[cigar_off[i] .. cigar_off[i+1]) and
 * letters seq_chars[seq_off[i] .. seq_off[i+1]) (no terminators) -- for hosts that keep their reads in arenas (and for
 * bindings where an array of C strings is expensive to build). */
int rambl_batch_add_subgroup_packed(rambl_batch* b, const char* gene, int32_t n_reads, const int32_t* pos,
                                    const int64_t* cigar_off, const char* cigar_chars, const int64_t* seq_off,
                                    const char* seq_chars, const int32_t* copies, const int32_t* pair_off,
                                    const int32_t* pair_val);

/* A subgroup whose graph was built elsewhere (e.g. by the reference's own PartialOrderGraph): the `nodes`
 * vector flattened in order.  Node u has AlignState state[u] (mat=0, mis, ins, del; PartialOrderGraph.hpp:82),
 * label label_chars[label_off[u]..label_off[u+1]), ordered successors out_to[out_off[u]..out_off[u+1]) and an
 * ordered read pool: entry e has read id pool_rid[e], copy number pool_copies[e] and letters
 * pool_chars[pool_str_off[e]..pool_str_off[e+1]).  Node 0 must be "^", exactly one node "$".  read_copies[r] is
 * the copy number of unique read r; pairs as in rambl_batch_add_subgroup.  The reads-over-edge counts
 * (number_of_reads_cover_nodes, PartialOrderGraph.cpp:1218-1244) are derived here.  Such a subgroup is ready
 * for rambl_batch_infer at once.  Returns the subgroup index (>= 0) or -error. */
int rambl_batch_add_graph(rambl_batch* b, int32_t n_nodes, int32_t n_reads, const uint8_t* state,
                          const int32_t* label_off, const char* label_chars, const int32_t* out_off,
                          const int32_t* out_to, const int32_t* pool_off, const int32_t* pool_rid,
                          const int32_t* pool_copies, const int32_t* pool_str_off, const char* pool_chars,
                          const int32_t* read_copies, const int32_t* pair_off, const int32_t* pair_val);

/* PartialOrderGraph::build (PartialOrderGraph.cpp:67-265) for every subgroup added so far; the
 * insertion alignments of all subgroups run in one device launch. */
int rambl_batch_build_graphs(rambl_batch* b);

/* The same construction with the device step taken out, for hosts that already hold the aligned
 * rows: thread_reads() splices the reads and lists the alignment problems (text: one line
 * "P <index> <n>" per problem followed by its n sequences, one per line);
 * finish_graphs_with_rows() takes the rows in the same layout ("P <index> <n>", then n rows). */
int rambl_batch_thread_reads(rambl_batch* b);
char* rambl_batch_msa_problems_text(rambl_batch* b);
int rambl_batch_finish_graphs_with_rows(rambl_batch* b, const char* rows_text);

/* PartialOrderGraph::infer_strains(strains, read_pairs, n, e, tau, diff) (NonparametricClustering.cpp:704-708)
 * and, when do_assign != 0, PartialOrderGraph::read_assign(strains, reads, read_pairs, n)
 * (NonparametricClustering.cpp:776-836) followed by main()'s abundance sort (StrainCall.cpp:1027),
 * for every subgroup.  Subgroups whose candidates were all pruned get status RAMBL_ERR_NO_STRAINS, and
 * subgroups whose candidate set outgrew the kernels (more than 256 live strains, which needs
 * degenerate abundances) get RAMBL_ERR_CAPACITY; the call itself still returns RAMBL_OK. */
int rambl_batch_infer(rambl_batch* b, int32_t n, float e, float tau, float diff, int32_t do_assign, int32_t keep_loglik);

/* rambl_batch_build_graphs + rambl_batch_infer in one call that overlaps them: a batch of more than one wave of the walk
 * kernel (one subgroup per SM) is dealt into a first chunk of one wave and the rest, and the host builds the graphs of the
 * second chunk while the device searches the strains of the first (two CUDA streams, two driver threads).
 * Same results as the two calls (subgroups never interact); what StrainCall's main() loop is to the CLI. */
int rambl_batch_solve(rambl_batch* b, int32_t n, float e, float tau, float diff, int32_t do_assign, int32_t keep_loglik);
/* The chunk layout rambl_batch_solve would use for n_subgroups on a device with `sms` SMs: writes up to `cap` chunk bounds
 * (chunk c = subgroups bounds[c] .. bounds[c+1]-1) and returns how many there are (chunks + 1), or a negative error.  Needs no
 * device; honours $RAMBL_SOLVE_FIRST / $RAMBL_SOLVE_CHUNKS. */
int32_t rambl_solve_layout(int32_t n_subgroups, int32_t sms, int32_t* bounds, int32_t cap);

/* What the device-resident walk will do with subgroup sg (graphs must be built; needs no device -- the level tables are a
 * property of the graph): out = { eligible, stops at an early "$" and hands over to the level-synchronous path, reason when not
 * eligible (engine.hpp), graph levels, read-pool entries over all levels, most entries of one level, most draws of one level,
 * levels on which a one-letter strain label can meet a multi-letter read string (deviation (i)) }. */
int rambl_batch_walk_plan(const rambl_batch* b, int32_t sg, int64_t out[8]);

/* ---- results */
int32_t rambl_batch_num_subgroups(const rambl_batch* b);
int32_t rambl_batch_num_nodes(const rambl_batch* b, int32_t sg);
/* what = 0: one NODE line per node (id, state, label, level, ordered OUT / IN lists, ordered read pool);
 * what = 1: PartialOrderGraph::output_edge (PartialOrderGraph.cpp:318-337), the -G output of StrainCall */
char* rambl_batch_graph_text(const rambl_batch* b, int32_t sg, int32_t what);
int32_t rambl_batch_status(const rambl_batch* b, int32_t sg);
int32_t rambl_batch_num_strains(const rambl_batch* b, int32_t sg);
/* strain k in the order streaming_clustering leaves them; order[] lists them by final abundance */
int rambl_batch_strain(const rambl_batch* b, int32_t sg, int32_t k, double* abundance_infer, double* abundance,
                       int32_t* path_len);
int rambl_batch_strain_path(const rambl_batch* b, int32_t sg, int32_t k, int32_t* path);
char* rambl_batch_strain_sequence(const rambl_batch* b, int32_t sg, int32_t k, int32_t plain); /* strain_seq / plain_seq */
int rambl_batch_strain_sub(const rambl_batch* b, int32_t sg, int32_t k, double* sub36);
int rambl_batch_strain_loglik(const rambl_batch* b, int32_t sg, int32_t k, double* loglik, int32_t n_reads);
int rambl_batch_order(const rambl_batch* b, int32_t sg, int32_t* order);
/* all of the above as text: STAGE/STRAIN/PATH/SEQ/PLAIN/SUB/LOGLIK lines (same layout as the oracle's dumps) */
char* rambl_batch_strains_text(const rambl_batch* b, int32_t sg);
/* the FASTA records StrainCall prints for this subgroup (StrainCall.cpp:1032-1046) */
char* rambl_batch_fasta(const rambl_batch* b, int32_t sg, const char* gene_name, int32_t p0, int32_t p1, float tau);
int rambl_batch_stats(const rambl_batch* b, rambl_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* RAMBL_B200_H */
