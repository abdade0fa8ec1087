/*
 * swb200.h -- C ABI of the B200-native Smith-Waterman fill / backtrack path.
 *
 * The reference (chunhualiao/Smith-Waterman) has no library interface; its hot
 * path is reached through
 *   (1) the per-cell helpers  similarityScore(i,j,H,P,&maxPos) / backtrack(P,maxPos)
 *       driven by the nDiag wavefront loop in main()      (omp_smithW.c:52-54,203-216,331-420)
 *   (2) the whole-fill operator of the rotated variants
 *       smithWaterman(a,b,w,h,H,P,&maxloc)                 (rotated-cuda/sw-rotated-cuda-unified.cu:198-215,
 *                                                           rotated-cuda/sw-rotated-omp.cc:192-209)
 *   (3) the CLI  ./omp_smithW <number_of_col> <number_of_rows>   (omp_smithW.c:8,91-96)
 * Every entry point below names the reference interface it replaces.  All of
 * them return 0 on success and a negative swb_status otherwise (the reference
 * prints and exit(0)s on CUDA errors, simple-cuda/sw-default-discrete.cu:101-108;
 * we never exit).  Calls on different streams/devices are independent; the only
 * process-wide state is, per device, one private CUDA memory pool for the per-call
 * workspace (it keeps at most 2 GB of freed blocks; the default pool is not touched)
 * and two streams that swb_fill_pairs_async forks large pairs onto (created on
 * first use).  There is NO CPU fallback: without a CUDA device every compute
 * entry point returns SWB_ERR_CUDA.
 *
 * Data contract (identical to the reference, omp_smithW.c:109-118,336):
 *   a: m bytes (columns), b: n bytes (rows)
 *   H, P: (n+1) x pitch int32, row-major, pitch >= m+1 (reference: pitch == m+1),
 *         row 0 and column 0 are zero and ARE written by the fill (the caller's
 *         buffers need not be initialised, as in (2))
 *   P codes NONE 0, UP 1, LEFT 2, DIAGONAL 3 (omp_smithW.c:33-36); backtrack
 *   multiplies the path cells by -1 (omp_smithW.c:32,417)
 *   maxPos = pitch*i + j of the first cell, in the reference's scan order
 *   (anti-diagonal ascending, then row descending), that attains the global
 *   maximum; 0 if no score is positive (omp_smithW.c:173,384-387)
 */
#ifndef SWB200_H
#define SWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SWB_OK            =  0,
    SWB_ERR_ARG       = -1,   /* null pointer, non-positive size, pitch < m+1 ...          */
    SWB_ERR_ALIGN     = -2,   /* dH/dP not 16-byte aligned                                 */
    SWB_ERR_CUDA      = -3,   /* a CUDA call failed (see swb_last_cuda_error)              */
    SWB_ERR_RANGE     = -4,   /* sizes/scores outside what 32-bit packed scores can hold   */
    SWB_ERR_NOMEM     = -5,   /* device or pinned-host allocation failed                   */
    SWB_ERR_IO        = -6    /* sequence file / manifest unreadable or malformed          */
} swb_status;

/* omp_smithW.c:75-77 (matchScore, missmatchScore, gapScore); NULL means 3,-3,-2 */
typedef struct { int32_t match, mismatch, gap; } swb_scoring;

/* parameters.h:1-2 of the reference; kept for CLI / log compatibility only */
#define SWB_FACTOR 128
#define SWB_CUTOFF 1024

/* Tuning knobs (0 = library default).  Not part of the reference surface. */
typedef struct swb_timer swb_timer;   /* CUDA-event pair around the fill kernel only */
typedef struct {
    int32_t warps_per_band;   /* compute warps (strips: 96 rows for a single pair, 64 rows in a batch) per CTA band, 1..2 (larger values are clamped to 2) */
    int32_t reserved[5];
    swb_timer* timer;         /* if set, swb_fill_async brackets the fill kernel launch with its events */
    uint64_t* trace;          /* developer tool: DEVICE buffer of ceil(n/64)*8 uint64 globaltimer stamps (8 per strip; strips are 96 rows for a single pair, 64 in a batch), or NULL; honoured by the -DSWB_TRACE developer build only */
} swb_tuning;

/* Kernel-only timing for the roofline figure: events are recorded on the call's stream
 * immediately before/after the fill kernel launch (not the prep/argmax kernels). */
int  swb_timer_create(swb_timer** t, int device);
int  swb_timer_elapsed_ms(swb_timer* t, float* ms);   /* synchronises on the stop event */
void swb_timer_destroy(swb_timer* t);

const char* swb_strerror(int status);
const char* swb_last_cuda_error(void);          /* thread-local text of the last CUDA failure */
int  swb_version(void);
int  swb_device_count(void);

/* ---- device-pointer API (the kernel boundary) --------------------------------
 * Replaces the nDiag loop + similarityScore (omp_smithW.c:203-216,331-388) and
 * smithWaterman() of the rotated variants.
 *   a, b      : host OR device pointers (detected); m = cols, n = rows
 *   dH, dP    : caller-owned DEVICE buffers, (n+1)*pitch int32 each, 16-byte aligned
 *   d_maxPos  : DEVICE int64 (may be NULL), written asynchronously
 *   stream    : cudaStream_t (NULL = default stream); the call only enqueues work
 */
int swb_fill_async(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
                   int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream,
                   const swb_tuning* tuning);

/* Same, but synchronises the stream and returns maxPos / maxScore to the host. */
int swb_fill(const char* a, int64_t m, const char* b, int64_t n,
             const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
             int64_t* maxPos, int device, void* stream);

/* Replaces backtrack(P, maxPos) (omp_smithW.c:405-420): negates P along the path
 * in place on the device.  maxPos == 0 (no positive score) is a defined no-op
 * (the reference has undefined behaviour there).  d_maxPos, when non-NULL, is a
 * DEVICE pointer read on the stream (chain after swb_fill_async without a sync)
 * and takes precedence over maxPos.  d_pathLen: DEVICE int64 or NULL. */
int swb_backtrack_async(int32_t* dP, int64_t pitch, int64_t maxPos, const int64_t* d_maxPos,
                        int64_t* d_pathLen, int device, void* stream);
int swb_backtrack(int32_t* dP, int64_t pitch, int64_t maxPos, int64_t* path_len,
                  int device, void* stream);

/* ---- host-buffer API (what a caller of the reference program would bind) -----
 * One call = the timed region of simple-cuda/sw-default-discrete.cu:382-444
 * (H2D of a,b; fill; D2H of the matrices) plus backtrack.  H and P are HOST
 * buffers of (n+1)*(m+1) int32 (pitch m+1); either may be NULL to skip its
 * copy-back.  P is returned AFTER backtrack when do_backtrack != 0. */
int swb_align_host(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* H, int32_t* P,
                   int64_t* maxPos, int64_t* path_len, int do_backtrack, int device);

/* Reusable context for repeated host-buffer calls of one shape (keeps the device
 * matrices and workspace alive between calls; swb_align_host creates/destroys
 * one internally). */
typedef struct swb_ctx swb_ctx;
int  swb_ctx_create(swb_ctx** ctx, int64_t m, int64_t n, int device);
int  swb_ctx_align(swb_ctx* ctx, const char* a, const char* b, const swb_scoring* scoring,
                   int32_t* H, int32_t* P, int64_t* maxPos, int64_t* path_len, int do_backtrack);
int32_t* swb_ctx_dH(swb_ctx* ctx);
int32_t* swb_ctx_dP(swb_ctx* ctx);
void swb_ctx_destroy(swb_ctx* ctx);

/* Packed transfer of H and P (swb_pack.cu).  The reference's matrices are int32 in host memory
 * (omp_smithW.c:203-216), so a device fill ends with 8 bytes per cell crossing PCIe -- 70 times the
 * duration of the fill.  On the wire one byte per cell is enough: bits 7..3 = H[i][j] - H[i][j-1] + 16
 * (the recurrence bounds the row step of H by gap .. match - gap, omp_smithW.c:331-388), bits 2..0 =
 * P[i][j] + 3 (directions 0..3, negated on the path, omp_smithW.c:405-420).
 *   swb_pack_rows_async  DEVICE: rows row0 .. row0+nrows-1, columns 0 .. cols-1 of dH/dP (row pitch `pitch`) ->
 *                        d_packed (row pitch packed_pitch >= cols, a multiple of 4; swb_packed_pitch(cols) gives
 *                        the canonical one).  d_row_base == NULL: column -1 counts as 0 (a whole matrix: column 0
 *                        is 0).  d_row_base != NULL (DEVICE, nrows int32): receives H of column 0 of every row and
 *                        column 0's step is stored as 0 -- for sub-matrices whose first column is not 0 (a column
 *                        strip: pass dH + c0, dP + c0).  *d_overflow (device int, zeroed by the caller) is set to 1
 *                        when a value does not fit; the packed bytes are then meaningless and the caller must copy
 *                        the int32 matrices instead.
 *   swb_expand_rows      HOST: packed rows -> int32 H and/or P (either may be NULL), bit-exact, starting every row
 *                        from row_base[r] (NULL: 0), on `threads` host threads (0 = swb_host_threads():
 *                        SWB_HOST_THREADS, else the cores this process may run on).
 *   swb_d2h_packed       both halves for a whole (sub-)matrix: packs rows 0..nrows-1 x columns 0..cols-1 of dH/dP,
 *                        copies the bytes in chunks and expands them into HOST H / P (row pitch host_pitch; either
 *                        may be NULL) while later chunks are in flight; falls back to plain copies when the flag
 *                        is raised.  Synchronous (returns when H and P are complete).  d_scratch (device) and
 *                        h_scratch (pinned host) hold swb_d2h_packed_scratch_bytes(nrows, cols) bytes each.
 * swb_ctx_align uses it for matrices of 32 MiB and more (SWB_PACKED_D2H=0 turns that off, =1 forces it); the caller
 * still receives the reference's int32 matrices.  A caller that only needs parts of H/P (say the rows a path
 * crosses) can keep the packed form and expand just those rows. */
int64_t swb_packed_pitch(int64_t cols);
int  swb_host_threads(void);
int  swb_pack_rows_async(const int32_t* dH, const int32_t* dP, int64_t pitch, int64_t row0, int64_t nrows, int64_t cols,
                         unsigned char* d_packed, int64_t packed_pitch, int* d_overflow, int32_t* d_row_base, int device,
                         void* stream);
int  swb_expand_rows(const unsigned char* packed, int64_t packed_pitch, int64_t nrows, int64_t cols,
                     int32_t* H, int32_t* P, int64_t pitch, const int32_t* row_base, int threads);
size_t swb_d2h_packed_scratch_bytes(int64_t nrows, int64_t cols);
int  swb_d2h_packed(const int32_t* dH, const int32_t* dP, int64_t pitch, int64_t nrows, int64_t cols,
                    int32_t* H, int32_t* P, int64_t host_pitch, void* d_scratch, void* h_scratch, int threads, int device,
                    void* stream);

/* Score-only variant (no H/P stores): max score and maxPos with the reference
 * tie-break.  Replaces the -DSKIP_BACKTRACK style runs of the reference's
 * variants (omp_smithW-v1-refinedOrig.cpp:190-192) for callers that only need
 * the score. */
int swb_score_only(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* maxScore, int64_t* maxPos,
                   int device, void* stream);

/* Asynchronous / batched score-only form: npairs equally shaped pairs (a = npairs*m bytes,
 * b = npairs*n bytes, pair k at offset k*m / k*n), d_maxPos and d_maxScore are DEVICE arrays
 * of npairs entries (either may be NULL).  maxPos uses pitch m+1. */
int swb_score_only_async(const char* a, int64_t m, const char* b, int64_t n, int64_t npairs,
                         const swb_scoring* scoring, int64_t* d_maxPos, int32_t* d_maxScore,
                         int device, void* stream, const swb_tuning* tuning);

/* Batch of npairs independent, equally shaped pairs in ONE launch (BASELINE config "batch of
 * 65536 independent 256x256 pairs"; the reference runs one pair per process).  Sequences are
 * packed: a = npairs*m bytes, b = npairs*n bytes.  Pair k's matrices start at dH/dP +
 * k*pair_stride int32 (pair_stride >= (n+1)*pitch, a multiple of 4 keeps every pair 16-byte
 * aligned); each has the single-pair layout of swb_fill_async.  d_maxPos / d_maxScore: DEVICE
 * arrays of npairs entries (maxPos relative to the pair's own matrix) or NULL. */
int swb_fill_batch_async(const char* a, int64_t m, const char* b, int64_t n, int64_t npairs,
                         const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch, int64_t pair_stride,
                         int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream, const swb_tuning* tuning);

/* Batch of independent pairs of ANY shapes (SURVEY 8(b) swb_fill_batch with per-pair m[], n[]; the reference
 * runs one pair per process).  HOST arrays of npairs entries: a_off / b_off = byte offset of pair k's sequences in
 * a / b (host or device concatenations), m / n = its lengths, hp_off = int32 offset of its matrices in dH / dP (a
 * multiple of 4; the matrices have the single-pair layout with pitch m[k]+1).  Runs of consecutive pairs with the
 * same shape, packed sequences and a uniform matrix stride go out as ONE batched launch (the 65536 x 256x256
 * configuration is a single run); other pairs get a launch each.  LARGE pairs (n >= 14208 rows and m >= 4096: one
 * pair alone covers half of the GPU's strip slots) always run one by one on the single-pair kernel, alternating
 * between two internal streams that fork from and join into `stream`: the wavefront of pair k+1 starts on the SMs
 * the ramp-down of pair k leaves idle (45000 x 45000: 4.60 -> 3.66 ms per pair).  The call stays asynchronous.
 * d_maxPos / d_maxScore: DEVICE arrays of npairs entries (maxPos relative to the pair's own matrix) or NULL. */
int swb_fill_pairs_async(const char* a, const int64_t* a_off, const int64_t* m,
                         const char* b, const int64_t* b_off, const int64_t* n,
                         const int64_t* hp_off, int64_t npairs, const swb_scoring* scoring,
                         int32_t* dH, int32_t* dP, int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream);

/* Pair-wise sharding of a batch over the GPUs of a box (BASELINE config "batch of 65536 independent 256x256 pairs
 * ... sharded pair-wise across 8 B200"; SURVEY 8(e): contiguous blocks, no exchange): shard `shard` of `nshards`
 * owns pairs first .. first+count-1.  Every GPU then calls swb_fill_batch_async / swb_fill_pairs_async on its own
 * pairs; nothing crosses the GPUs. */
int swb_shard_pairs(int64_t npairs, int nshards, int shard, int64_t* first, int64_t* count);

/* ---- column-strip mode: ONE pair across several GPUs (BASELINE config "100000x100000 single pair,
 * column-strip wavefront pipelined across 2/4/8 B200 with NVLink P2P boundary exchange").
 * The reference has no multi-GPU form; this extends the nDiag wavefront (omp_smithW.c:203-216).
 * GPU g owns the columns col0+1 .. col0+m_local of every row and holds them as its own row-major
 * matrices dH/dP of (n+1) x pitch, pitch >= m_local+1; local column j is global column col0+j and
 * local column 0 is a copy of the last column of the GPU on the left (the zero column for GPU 0).
 *   left_in     DEVICE int32[n+1] on this GPU: H of that column, written by the left GPU's fill
 *               kernel with peer stores while it runs (NULL for the first strip)
 *   left_flags  DEVICE int32[swb_strip_flag_count(n)] on this GPU: flag k == epoch once the rows of
 *               piece k are in left_in (NULL with left_in)
 *   right_out / right_flags  the right GPU's left_in / left_flags as PEER pointers valid on this GPU
 *               (cudaIpcOpenMemHandle / peer access; NULL for the last strip)
 *   epoch       nonzero, different from the previous call's on the same buffers
 * The fill kernel of strip g+1 waits on the flags piece by piece, so all strips run concurrently,
 * each a few row-blocks behind its left neighbour; there is no collective in the data path.
 * Ordering rules for the caller: (1) strips that share a device must be enqueued left to right on ONE stream;
 * (2) the boundary buffers of a call may be reused only after the right neighbour has finished the call that
 * read them (swb_multi_* synchronises every stream per call; strips.StripPipeline separates calls by the maxPos
 * all-gather or a barrier) -- alternate two buffer sets by call parity to let the left GPU start early;
 * (3) a strip whose left neighbour never publishes its flags traps after ~10 s of waiting (the launch fails
 * with a CUDA error) instead of hanging.
 * d_maxPos / d_maxScore: the LOCAL maximum over local columns 1..m_local (index i*pitch + j_local,
 * reference tie-break).  In P, local column 0 of a strip with a left neighbour holds 5 (not a
 * reference code): the backtrack hand-off marker. */
int swb_fill_strip_async(const char* a_local, int64_t m_local, const char* b, int64_t n,
                         const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
                         const int32_t* left_in, const int32_t* left_flags, int32_t* right_out, int32_t* right_flags,
                         int32_t epoch, int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream,
                         const swb_tuning* tuning);
int64_t swb_strip_flag_count(int64_t n);

/* backtrack (omp_smithW.c:405-420) from an arbitrary start cell of a strip: follows P from startPos,
 * negating the path, until a NONE cell or the hand-off marker of local column 0.  d_endPos (DEVICE
 * int64) receives the index of the cell that ended the walk: if it lies in local column 0 of a strip
 * with a left neighbour the path continues at the same row in the last column of that neighbour. */
int swb_backtrack_from_async(int32_t* dP, int64_t pitch, int64_t startPos, int64_t* d_pathLen, int64_t* d_endPos,
                             int device, void* stream);

/* ---- C++ host driver for the column-strip mode inside ONE process (SURVEY 8(b) swb_fill_multi): the pair is
 * split into ndev contiguous column blocks, strip g on devices[g] (devices may repeat: such strips share a stream
 * and run left to right), peer access is enabled between neighbours, every strip has its own stream and all fill
 * kernels run concurrently, linked only by the NVLink boundary stores of swb_fill_strip_async.  maxPos is reduced
 * on the host from one (score, position) pair per GPU with the reference's tie-break (omp_smithW.c:384-387) and
 * the backtrack (omp_smithW.c:405-420) walks right to left over the strips.  a, b: HOST sequences of the whole pair.
 * maxPos indexes the (n+1) x (m+1) matrix of the whole pair.  swb_multi_gather_host copies the strips into the
 * reference's row-major host layout (pitch m+1; either pointer may be NULL); swb_multi_strip exposes one strip's
 * device slab (local column j = global column col0 + j; local column 0 of strips g > 0 holds the left
 * neighbour's last column in H and the hand-off marker 5 in P). */
typedef struct swb_multi swb_multi;
int  swb_multi_create(swb_multi** out, int64_t m, int64_t n, const int* devices, int ndev);
int  swb_multi_fill(swb_multi* h, const char* a, const char* b, const swb_scoring* scoring,
                    int64_t* maxPos, int32_t* maxScore);
int  swb_multi_backtrack(swb_multi* h, int64_t maxPos, int64_t* path_len);
int  swb_multi_align(swb_multi* h, const char* a, const char* b, const swb_scoring* scoring,
                     int64_t* maxPos, int32_t* maxScore, int64_t* path_len, int do_backtrack);  /* = fill [+ backtrack] */
int  swb_multi_strips(const swb_multi* h);
int  swb_multi_strip(const swb_multi* h, int g, int* device, int64_t* col0, int64_t* m_local, int64_t* pitch,
                     int32_t** dH, int32_t** dP);
int  swb_multi_gather_host(swb_multi* h, int32_t* H, int32_t* P);
void swb_multi_destroy(swb_multi* h);
/* one-shot form: create, align, gather (H, P: HOST (n+1) x (m+1) int32 or NULL), destroy */
int  swb_fill_multi(const char* a, int64_t m, const char* b, int64_t n, const swb_scoring* scoring,
                    const int* devices, int ndev, int32_t* H, int32_t* P,
                    int64_t* maxPos, int64_t* path_len, int do_backtrack);

/* ---- alignment emission: the step right after backtrack(P, maxPos) (omp_smithW.c:405-420), which
 * only negates the path.  swb_traceback_async is swb_backtrack_from_async that also writes the moves of the path
 * to d_moves (DEVICE, one byte per path cell, capacity >= n + m; may be NULL), in walk order = from the start cell
 * backwards: 1 UP, 2 LEFT, 3 DIAGONAL (the P codes, omp_smithW.c:33-36).  d_startPos (DEVICE int64, e.g. the d_maxPos
 * of swb_fill_async) takes precedence over startPos.  The host helpers turn moves into a CIGAR string (sequence
 * order; a = columns is the reference, b = rows the query: DIAGONAL -> M, UP -> I, LEFT -> D; returns the length the
 * string needs, writes at most cap-1 characters + NUL) and into the two gapped strings (out_a, out_b: n+1 bytes). */
int swb_traceback_async(int32_t* dP, int64_t pitch, int64_t startPos, const int64_t* d_startPos, int64_t* d_pathLen,
                        int64_t* d_endPos, unsigned char* d_moves, int device, void* stream);
int64_t swb_cigar_from_moves(const unsigned char* moves, int64_t n, char* cigar, size_t cap);
int swb_alignment_from_moves(const unsigned char* moves, int64_t n, const char* a, const char* b, int64_t startPos,
                             int64_t pitch, char* out_a, char* out_b);

/* Boundary buffers that another process maps (cudaMalloc + CUDA IPC): allocation (zero-filled),
 * 64-byte IPC handle, mapping a peer's handle on `device`, and peer access between two devices of
 * one process. */
void* swb_ipc_alloc(size_t bytes, int device);
void  swb_ipc_free(void* p, int device);
int   swb_ipc_get_handle(void* p, unsigned char* out64);
int   swb_ipc_open(const unsigned char* handle64, int device, void** out);
int   swb_ipc_close(void* p, int device);
int   swb_enable_peer(int device, int peer);

/* ---- real sequence input, host only: replaces generate() (omp_smithW.c:489-519) as the source of
 * a and b.  Files are FASTA ('>' headers; letters are upper-cased, white space dropped; a file without a header
 * is one anonymous record) or UCSC .2bit (detected by its signature; N blocks come back as 'N').  swb_seq_read
 * returns record `record` (0-based) as a malloc'ed, NUL-terminated string (free with swb_seq_free); name may be NULL.
 * A manifest is a text file with one pair per line, "<fileA>[:record] <fileB>[:record]" ('#' comments, paths
 * relative to the manifest): the pairs of a variable-length batch for swb_fill_pairs_async. */
int  swb_seq_count(const char* path, int64_t* nrecords);
int  swb_seq_read(const char* path, int64_t record, char** seq, int64_t* len, char* name, size_t name_cap);
void swb_seq_free(char* seq);
typedef struct swb_manifest swb_manifest;
int     swb_manifest_load(const char* path, swb_manifest** out);
int64_t swb_manifest_pairs(const swb_manifest* m);
int     swb_manifest_pair(const swb_manifest* m, int64_t k, const char** a, int64_t* alen, const char** b, int64_t* blen);
void    swb_manifest_free(swb_manifest* m);

/* The reference's generate() (omp_smithW.c:489-519): srand(seed) then m+1 draws
 * for a and n+1 draws for b with this libc's rand(), 0->A 2->C 3->G else T.
 * Host-side helper so that callers get the reference's sequences for a seed. */
void swb_generate(unsigned seed, int64_t m, int64_t n, char* a, char* b);

/* Pinned host memory for the host-buffer API (the D2H of H and P dominates an
 * end-to-end call; pageable buffers halve its rate).  NULL on failure. */
void* swb_host_alloc(size_t bytes);
void  swb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */
