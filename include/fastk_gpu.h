/*  fastk_gpu.h -- C ABI of libfastk_gpu.so, the B200 (sm_100a) replacement for FastK's k-mer counting
 *  hot path.  Plain C: opaque handle, plain pointers and sizes, no torch / C++ types.
 *
 *  What it replaces in the reference (paths relative to the FastK source tree):
 *     split.c    Distribute_Block / Split_Kmers   (split.c:1016-1393, 1407-1713)   -> fkgpu_ingest
 *     count.c    Sorting                          (count.c:1202-1914)              -> fkgpu_finish
 *     MSDsort.c  Supermer_Sort / Weighted_Kmer_Sort (MSDsort.c:458-544)            -> kernels behind fkgpu_finish
 *     LSDsort.c  LSD_Sort                         (LSDsort.c:115-271)              -> kernels behind fkgpu_finish
 *     merge.c    Merge_Profiles                   (merge.c:761-1006)               -> fkgpu_profiles
 *  What stays on the host and calls this library: io.c (Scan_All_Input -> Distribute_Block callback,
 *  io.c:2659-2699), table.c (Merge_Tables consumes [KMER_BYTES key][u16 count] runs, table.c:382-394),
 *  libfastk.c (file readers) and the FastK.c driver.  INTEGRATION.md shows the shim.
 *
 *  Error convention: every entry point returns 0 on success, a negative FKGPU_E_* code otherwise;
 *  fkgpu_last_error() gives the message.  The reference's convention (message on stderr, then
 *  Clean_Exit(1), FastK.c:181-221) is applied by the host shim, not in here.
 *
 *  There is NO CPU fallback: without a CUDA device fkgpu_create fails with FKGPU_E_NODEVICE.
 */
#ifndef FASTK_GPU_H
#define FASTK_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FKGPU_OK             0
#define FKGPU_E_NODEVICE    -1    /* no CUDA device / driver                                  */
#define FKGPU_E_CUDA        -2    /* a CUDA runtime call or kernel failed                     */
#define FKGPU_E_ARG         -3    /* bad argument (k out of range, NULL pointer, bad tid ...)  */
#define FKGPU_E_NOMEM       -4    /* host or device allocation failed                         */
#define FKGPU_E_STATE       -5    /* call out of order (e.g. ingest after finish)             */
#define FKGPU_E_UNSUPPORTED -6    /* k > FKGPU_MAX_K                                          */

#define FKGPU_MAX_K        64     /* packed key = 1 or 2 64-bit words                         */
#define FKGPU_HIST_BINS    32768  /* bins 1..32767 used; 32767 = ">= 32767" (MSDsort.c:498)   */

typedef struct fkgpu_ctx fkgpu_ctx;

/*  Mirrors the option globals the replaced stages read (FastK.h:34-83). */
typedef struct
  { int32_t  kmer;          /* KMER        -k          1 .. FKGPU_MAX_K                         */
    int32_t  do_table;      /* DO_TABLE    -t<cutoff>  0 = no table, else emit counts >= cutoff */
    int32_t  do_profile;    /* DO_PROFILE  -p          keep reads on the device for fkgpu_profiles */
    int32_t  bc_prefix;     /* BC_PREFIX   -bc<n>      ignore first n bases of every read       */
    int32_t  device;        /* CUDA device ordinal                                             */
    int32_t  nthreads;      /* # of distinct tid values that will call fkgpu_ingest (ITHREADS)  */
    int64_t  reserve_bases; /* hint: total bases expected (0 = grow as needed)                 */
    int64_t  mem_limit;     /* SORT_MEMORY -M<GB> in bytes: most device memory the working buffers of a count may take
                               (0 = whatever is free).  An input whose one-pass working set does not fit is counted in
                               several ROUNDS over disjoint minimizer-bucket ranges, each leaving one sorted run -- the
                               role of NPARTS in the reference (FastK.c:419-429, count.c:1337, merged by table.c:240-313) */
  } fkgpu_config;

/*  Result of fkgpu_finish / fkgpu_count_packed.  All pointers are owned by the context and stay
 *  valid until the next finish/reset/destroy.  table = ntable records of (kmer_bytes + 2) bytes:
 *  [kmer_bytes big-endian 2-bit packed canonical k-mer, unused low bits 0][u16 LE count saturated at 32767],
 *  strictly increasing by key -- exactly the record stream table.c:382-394 reads from its L-files.      */
typedef struct
  { int32_t        kmer;
    int32_t        kmer_bytes;        /* (2k+7)>>3, FastK.c:419                                  */
    int64_t        nbases;            /* bases ingested (terminators excluded)                   */
    int64_t        nreads;            /* reads ingested                                          */
    int64_t        nkmers;            /* valid k-mer instances counted                           */
    int64_t        ndistinct;         /* distinct canonical k-mers                               */
    const int64_t *hist;              /* host, [FKGPU_HIST_BINS]; hist[c] = # distinct k-mers with count c */
    int64_t        max_inst;          /* sum of true counts of k-mers with count >= 32767        */
    int64_t        ntable;            /* # of table records (count >= do_table); 0 if !do_table  */
    const uint8_t *table;             /* host (pinned) copy of the table records, NULL if !do_table or not fetched */
    const uint8_t *table_dev;         /* device copy of the same records                         */
    float          ms_pack;           /* device time of the stages, CUDA events, for -v reporting */
    float          ms_count;
    float          ms_total;
    /*  The table as sorted runs.  One run (nruns == 1, run_table[0] == table) unless the count took several rounds; then
     *  every round leaves one run, strictly increasing by key, the runs hold DISJOINT key sets (a canonical k-mer lives in
     *  one minimizer bucket), ntable is their total and table / table_dev are NULL: the consumer merges them, as
     *  Merge_Tables does with the NPARTS part files (table.c:382-394).                                               */
    int32_t        nruns;
    const int64_t *run_ntable;        /* [nruns] records of every run                              */
    const uint8_t *const *run_table;  /* [nruns] host (pinned) pointer of every run, NULL entries if not fetched */
  } fkgpu_result;

/* ---- life cycle ---------------------------------------------------------------------------------- */

int  fkgpu_create (const fkgpu_config *cfg, fkgpu_ctx **out);
void fkgpu_destroy(fkgpu_ctx *ctx);
int  fkgpu_reset  (fkgpu_ctx *ctx);             /* forget ingested reads and results, keep buffers */
const char *fkgpu_last_error(void);              /* thread-local message of the last failing call  */
int  fkgpu_device_count(void);                   /* # of CUDA devices, 0 if none                   */

/* ---- host-buffer path: what the reference's io.c callback binds to --------------------------------
 *  fkgpu_ingest replaces  void Distribute_Block(DATA_BLOCK *block, int tid)  (FastK.h:123):
 *    bases  = block->bases   concatenation of 0-terminated reads (FastK.h:92)
 *    boff   = block->boff    read i is bases+boff[i], boff[nreads] = total bytes (FastK.h:93)
 *    nreads = block->nreads
 *    rem    = block->rem     > 0: the last read continues in this tid's next block with a k-1 overlap
 *  Thread-safe for distinct tid; the block is copied before return (io.c reuses it immediately,
 *  io.c:552,565).  Reads of one tid keep their order; global read order is tid-major.             */
int  fkgpu_ingest(fkgpu_ctx *ctx, int tid, const char *bases, const int32_t *boff, int32_t nreads, int32_t rem);

/*  fkgpu_finish replaces  void Sorting(char *path, char *root)  (count.c:1202): uploads what was
 *  ingested, runs the device pipeline, fills res.  fetch_table != 0 also copies the table to
 *  pinned host memory (res->table).                                                               */
int  fkgpu_finish(fkgpu_ctx *ctx, int fetch_table, fkgpu_result *res);

/*  Count profiles (replaces count.c:817-1181 + merge.c:761-1006).  Valid after fkgpu_finish when
 *  do_profile was set.  For global read r (tid-major order), prof[off[r] .. off[r+1]) are the
 *  counts (saturated at 32767, 0 where the k-mer covers a non-acgt base) at each k-mer start of
 *  the read after its bc prefix; a read shorter than k has an empty range.  Pointers are pinned
 *  host memory owned by the context.                                                             */
int  fkgpu_profiles(fkgpu_ctx *ctx, int64_t *nreads, const int64_t **off, const uint16_t **prof);

/*  Relative profiles, -p:<table> (FastK.c:269-281; replaces Split_Table split.c:1943-2131 and the merge join of
 *  count.c:675-792).  records = n table records [kmer_bytes key][u16 LE count] in increasing key order, as read back from an
 *  existing .ktab (any cutoff).  Call it on a do_profile context BEFORE fkgpu_finish: finish then only packs the reads --
 *  nothing is counted, no histogram, no table (FastK.c:328-337) -- and fkgpu_profiles reports at every position the count
 *  the loaded table holds for the canonical k-mer there, 0 if it is absent.                                          */
int  fkgpu_load_profile_table(fkgpu_ctx *ctx, const uint8_t *records, int64_t n);

/*  GPU Fastmerge (Fastmerge.c:168-450): merges ntab k-mer tables of the same k, each n[t] records [kmer_bytes key][u16 LE count]
 *  in increasing key order (host memory), into one: the counts of equal k-mers are added and saturate at 32767; res->hist is
 *  the histogram of the merged counts; res->max_inst holds only the instances that unsaturated members of saturated sums
 *  stood for (Fastmerge.c:321-327) -- add the max_inst of the input histograms to it (Fastmerge.c:1009).  The context must
 *  have been created with do_table >= 1, no do_profile, and the tables' k.                                          */
int  fkgpu_merge_tables(fkgpu_ctx *ctx, const uint8_t *const *tables, const int64_t *n, int ntab, int fetch_table, fkgpu_result *res);

/*  # of whole reads each tid delivered (continuation pieces of a split read are not counted twice); part t+1 of
 *  the .prof output holds the reads of tid t (merge.c:926-928).  per_tid has cfg.nthreads entries.          */
int  fkgpu_read_counts(fkgpu_ctx *ctx, int64_t *per_tid);

/* ---- device-resident path (bench "value", multi-GPU stages) --------------------------------------
 *  Packed read stream: position i of the concatenated reads (one terminator position between reads)
 *    seq word i>>4 , bits 31-2*(i&15) .. 30-2*(i&15)   = base code a,c,g,t = 0..3
 *    val word i>>5 , bit  31-(i&31)                    = 1 iff position i holds an acgt base
 *  Both arrays must be followed by FKGPU_PACK_PAD zero words.                                      */
#define FKGPU_PACK_PAD 16

/*  ASCII (device or host memory) -> packed stream on the device; d_seq/d_val sized by fkgpu_packed_words. */
void fkgpu_packed_words(int64_t npos, int64_t *seq_words, int64_t *val_words);
int  fkgpu_pack_ascii_dev(fkgpu_ctx *ctx, const char *d_ascii, int64_t npos, uint32_t *d_seq, uint32_t *d_val);

/*  Whole counting pipeline over a packed stream already resident in HBM.  */
int  fkgpu_count_packed(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                        int fetch_table, fkgpu_result *res);

/*  Count profiles over a packed stream counted by fkgpu_count_packed on a context with do_profile (the device-resident
 *  form of fkgpu_profiles; count.c:817-1181): read i occupies positions [read_start[i], read_start[i] + read_len[i]) of
 *  the stream (host arrays).  Outputs as fkgpu_profiles.  Reads given in stream order without overlap take the fast
 *  path (counts written in output order, copied to the host in slices under the lookups); any other order is served
 *  through a per-position array and a gather.                                                                     */
int  fkgpu_profiles_packed(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                           const int64_t *read_start, const int32_t *read_len, int64_t nreads_in,
                           int64_t *nreads, const int64_t **off, const uint16_t **prof);

/*  Multi-GPU stages (one process per GPU; the exchange itself is done by the caller, e.g. NCCL
 *  all-to-all via torch.distributed -- see fastk_b200/multigpu.py):
 *   1. fkgpu_prefix_hist:   histogram of the top `bits` key bits of every valid canonical k-mer
 *                           (d_hist: 2^bits uint64 on the device, overwritten).
 *   2. fkgpu_scatter_prefix: writes every canonical k-mer as a 16-byte (k > 32) or 8-byte record,
 *                           grouped by that prefix, into d_records (capacity cap_records); d_offsets
 *                           (2^bits + 1 uint64, device) receives the group starts.
 *   3. fkgpu_count_records: sort / count / histogram / table over an arbitrary record array (e.g. the
 *                           records received from the peers).  d_records is CONSUMED: it is reused as the second
 *                           sort buffer and must have room for nrecords + 4 records.                            */
int  fkgpu_record_bytes(int kmer);
int  fkgpu_prefix_hist(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                       int bits, uint64_t *d_hist);
int  fkgpu_scatter_prefix(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                          int bits, const uint64_t *d_hist, void *d_records, int64_t cap_records,
                          uint64_t *d_offsets);
int  fkgpu_count_records(fkgpu_ctx *ctx, void *d_records, int64_t nrecords, int fetch_table, fkgpu_result *res);

/*  Multi-GPU stages of the super-mer path (k in 18..64; fastk_b200/multigpu.py drives them, one process per GPU).
 *  A super-mer record is 8 bytes, [minimizer bucket : <= 24][# k-mers - 1 : 6][strand of the minimizer : 1][GLOBAL position of its first base : >= 32],
 *  where the global position space is the concatenation of every rank's packed read stream (rank r starts at
 *  pos_base[r]).  Ranks own contiguous bucket ranges; ONE all-to-all moves the 8-byte records (not the k-mers), and the
 *  counting kernel of the owner gathers the bases straight from the source rank's HBM over NVLink (peer pointers obtained
 *  with the IPC calls below).  Every instance of a canonical k-mer falls in one bucket, so counts never merge across
 *  ranks; only the (much smaller) distinct (key | count) entries take a second all-to-all, by key prefix, when a sorted
 *  table is wanted.
 *   fkgpu_reads_alloc       packed-read buffers owned by the context (plain cudaMalloc: exportable over CUDA IPC)
 *   fkgpu_ipc_export/open/close   64-byte handle of a device allocation <-> pointer valid in this process
 *   fkgpu_super_bucket_bits bucket-id width every rank derives from the GLOBAL position count
 *   fkgpu_super_scan        reads -> records, partitioned by the top *hist_bits bucket bits; all outputs are device
 *                           pointers into context memory, valid until the next call on this context
 *   fkgpu_super_payload     the 32-byte left-aligned base string of every record, in record order, gathered from the rank's own
 *                           reads: exchanged beside the records (default; bulk NVLink transfers).  With d_payload == NULL in
 *                           fkgpu_super_count the bases are instead gathered from peer HBM inside the counting kernel
 *                           (measured: fine at 2 GPUs, collapses at 4 -- small random NVLink reads; FKGPU_MG=peer keeps it)
 *   fkgpu_super_count       received records (consumed; room for nrecords + 8) -> histogram / scalars in res and, if
 *                           want_entries, the distinct entries (fkgpu_entry_bytes(k) each: 16-byte key | count in the low 16 bits, or 24 bytes key, count) on the device
 *   fkgpu_entries_partition entries -> d_out ordered by the top `bits` key bits; d_hist [2^bits], d_offsets [2^bits+1]
 *   fkgpu_entries_sort      distinct entries (consumed; room for n + 4) -> key order -> table in res               */
#define FKGPU_IPC_HANDLE_BYTES 64
int  fkgpu_super_supported(int kmer);
int  fkgpu_entry_bytes(int kmer);            /* bytes of a distinct entry: 16 (key | count) up to k = 56, 24 (key, count) beyond */
int  fkgpu_super_bucket_bits(int kmer, int64_t npos_total);
int  fkgpu_reads_alloc(fkgpu_ctx *ctx, int64_t npos, uint32_t **d_seq, uint32_t **d_val);
int  fkgpu_ipc_export(fkgpu_ctx *ctx, const void *d_ptr, uint8_t *handle /*[FKGPU_IPC_HANDLE_BYTES]*/);
int  fkgpu_ipc_open  (fkgpu_ctx *ctx, const uint8_t *handle, void **d_ptr);
int  fkgpu_ipc_close (fkgpu_ctx *ctx, void *d_ptr);
int  fkgpu_super_scan(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos, int64_t npos_total,
                      int64_t pos_offset, const uint64_t **d_records, int64_t *nrecords, int64_t *nkmers,
                      const uint64_t **d_bucket_hist, const uint64_t **d_bucket_offsets, int32_t *hist_bits);
int  fkgpu_super_payload(fkgpu_ctx *ctx, const uint32_t *d_seq, int64_t pos_offset, int64_t npos_total,
                         const uint64_t *d_records, int64_t nrecords, void *d_payload /* nrecords x 32 bytes */);
int  fkgpu_super_count(fkgpu_ctx *ctx, uint64_t *d_records, int64_t nrecords, int64_t npos_total, int32_t nranks,
                       const uint32_t *const *seq_of_rank, const int64_t *pos_base, const void *d_payload,
                       void *payload_ready_event /* cudaEvent_t or NULL: the counting kernel waits for it, so the payload's
                                                    all-to-all can overlap the partition of the records */,
                       int want_entries, fkgpu_result *res, const void **d_entries, int64_t *nentries);
int  fkgpu_entries_partition(fkgpu_ctx *ctx, const void *d_entries, int64_t n, int bits, void *d_out,
                             uint64_t *d_hist, uint64_t *d_offsets);
int  fkgpu_entries_sort(fkgpu_ctx *ctx, void *d_entries, int64_t n, int fetch_table, fkgpu_result *res);

/* ---- multi-GPU count inside the library ---------------------------------------------
 *  One context per GPU (one process per GPU, or one thread per GPU of a single process), joined by one NCCL communicator
 *  (libnccl.so.2 is bound at run time; single-GPU users need no NCCL).  Reads are data-parallel: every rank ingests / packs
 *  ITS reads.  fkgpu_count_packed_multi (and fkgpu_finish on a context with a communicator) is COLLECTIVE: super-mer records
 *  and their base strings are exchanged so that every minimizer bucket is counted by one owner, the distinct entries are
 *  exchanged by key prefix and sorted by their owner.  Result per rank: res->hist / nkmers / ndistinct / max_inst are GLOBAL,
 *  res->table is this rank's key range; rank order == key order, so the global table is the rank-ordered concatenation
 *  (fkgpu_comm_info gives the sizes) -- what the reference gets from Merge_Tables (table.c:346-533).
 *   fkgpu_comm_id     a fresh communicator id (rank 0 calls it and hands the bytes to the others)
 *   fkgpu_comm_init   collective: joins rank `rank` of `nranks`
 *   fkgpu_comm_info   v[6] = nranks, rank, global table records, table offset of this rank, records sent, entries sent;
 *                     table_sizes (may be NULL) [nranks] records per rank                                                    */
#define FKGPU_COMM_ID_BYTES 128
int  fkgpu_comm_id  (uint8_t *id /*[FKGPU_COMM_ID_BYTES]*/);
int  fkgpu_comm_init(fkgpu_ctx *ctx, int nranks, int rank, const uint8_t *id);
int  fkgpu_comm_info(fkgpu_ctx *ctx, int64_t *v /*[6]*/, int64_t *table_sizes /*[nranks] or NULL*/);
int  fkgpu_count_packed_multi(fkgpu_ctx *ctx, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                              int fetch_table, fkgpu_result *res);

/*  Instrumentation for bench.py: # of kernel launches issued by this context so far, and the
 *  accumulated CUDA-event time / algorithmic bytes of the dominant kernel family (final sort+count). */
int64_t fkgpu_launch_count(fkgpu_ctx *ctx);
int     fkgpu_last_stats(fkgpu_ctx *ctx, int64_t *v /*[8]: path, super-mer records, distinct entries sorted, work groups, rounds,
                                                         hash classes split, k-mers of oversize buckets sent to the record pipeline,
                                                         super-mers expanded (the others were copies counted by weight)*/);
int     fkgpu_last_path(fkgpu_ctx *ctx);     /* which pipeline served the last count: 0 = 16-byte records, 1 = super-mers */
int     fkgpu_stage_times(fkgpu_ctx *ctx, float *ms /*[FKGPU_NSTAGES]*/, double *bytes /*[FKGPU_NSTAGES]*/);
#define FKGPU_NSTAGES 14
/* stage ids */
#define FKGPU_ST_PACK      0
#define FKGPU_ST_SCANHIST  1
#define FKGPU_ST_SCATTER   2
#define FKGPU_ST_L2HIST    3
#define FKGPU_ST_L2PART    4
#define FKGPU_ST_SORTCOUNT 5
#define FKGPU_ST_COMPACT   6
#define FKGPU_ST_PROFILE   7
/* super-mer path (k in 18..64): stages 5/6 then only order the distinct entries */
#define FKGPU_ST_SUPERSCAN 8     /* reads -> super-mer records                       */
#define FKGPU_ST_SUPERPART 9     /* bucket partition of the super-mer records        */
#define FKGPU_ST_BUCKET    10    /* on-chip expansion + hash count per bucket group  */
#define FKGPU_ST_ENTPART   11    /* prefix partition of the distinct (key|count) entries */
#define FKGPU_ST_SUPERREFINE 12   /* second partition level of the super-mer records  */
#define FKGPU_ST_SPILL     13    /* oversize buckets expanded to k-mer records and counted by the record pipeline */

#ifdef __cplusplus
}
#endif
#endif
