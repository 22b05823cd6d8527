/* zlb.h — thin C ABI between the host-side libzling mirror (C++ shim in libzling_b200/csrc/zl_api.cpp, or any
 * FFI: ctypes / cgo / JNI) and the sm_100a CUDA block pipeline.  Plain C types only; no exceptions cross this
 * boundary; every function returns 0 / a non-negative count on success and a negative zlb_status on failure,
 * with a human-readable message retrievable through zlb_last_error().  There is NO CPU fallback: when no CUDA
 * device (or no driver) is present zlb_create() fails with ZLB_E_NODEVICE.
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *   zlb_encoder_begin / zlb_encode_blocks / zlb_encoder_end
 *        the per-16-MiB-block body of baidu::zling::Encode()            src/libzling.cpp:187-284
 *        (ZlingRolzEncoder::Reset/Encode                                src/libzling_lz.cpp:128-316,
 *         frequency count + ZlingMakeLengthTable/ZlingMakeEncodeTable   src/libzling.cpp:219-229,
 *                                                                       src/libzling_huffman.cpp:41-138,
 *         nibble header + ZlingCodebuf bit packing + level feedback     src/libzling.cpp:232-266,
 *         framing flag/encpos/rlen/olen/payload/stop                    src/libzling.cpp:200,269-278)
 *   zlb_decoder_begin / zlb_decode_blocks / zlb_decoder_end
 *        the per-block body of baidu::zling::Decode()                   src/libzling.cpp:301-420
 *        (Huffman LUT decode :347-402, ZlingRolzDecoder::Reset/Decode   src/libzling_lz.cpp:318-399)
 *   zlb_encoder_get_state / zlb_encoder_set_state
 *        the state that outlives a block in the reference: m_mtf[256]   src/libzling_lz.h:105 (never reset) and
 *        current_level                                                  src/libzling.cpp:185,261-266
 * The stream-lifetime objects (zlb_encoder / zlb_decoder) exist because of that carried state: blocks of ONE
 * stream must be submitted in order; independent streams use independent encoder objects.
 */
#ifndef ZLB_H
#define ZLB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZLB_BLOCK_BYTES      16777216u  /* kBlockSizeIn       src/libzling.cpp:70 */
#define ZLB_SUBBLOCK_SYMBOLS 262144u    /* kBlockSizeRolz     src/libzling.cpp:71 */
#define ZLB_SUBBLOCK_BYTES   393216u    /* kBlockSizeHuffman  src/libzling.cpp:72 */
#define ZLB_STATE_BYTES      65540u     /* 256x256 MTF rank->byte tables + int32 current level */

typedef enum {
    ZLB_OK          = 0,
    ZLB_E_ARG       = -1,   /* bad argument (NULL, level outside 0..4, unaligned block submission ...) */
    ZLB_E_NODEVICE  = -2,   /* no CUDA device / driver: the product path refuses to run on the CPU */
    ZLB_E_CUDA      = -3,   /* a CUDA runtime call or kernel failed; see zlb_last_error() */
    ZLB_E_NOMEM     = -4,   /* host or device allocation failed */
    ZLB_E_OVERFLOW  = -5,   /* caller's output buffer too small */
    ZLB_E_FORMAT    = -6,   /* malformed compressed stream (where the reference throws std::runtime_error) */
    ZLB_E_NCCL      = -7    /* NCCL unavailable (libnccl.so.2 could not be loaded) or an NCCL call failed */
} zlb_status;

typedef struct zlb_ctx     zlb_ctx;      /* one per (process, GPU): device buffers, streams */
typedef struct zlb_encoder zlb_encoder;  /* one per stream being encoded */
typedef struct zlb_decoder zlb_decoder;  /* one per stream being decoded */

/* ---- device / context ------------------------------------------------------------------------------- */
int          zlb_device_count(void);                 /* number of CUDA devices, 0 if none, never fails */
zlb_ctx*     zlb_create(int device, int max_blocks); /* buffers for up to max_blocks 16 MiB blocks per call */
void         zlb_destroy(zlb_ctx* ctx);
int          zlb_max_blocks(const zlb_ctx* ctx);
const char*  zlb_last_error(void);                   /* thread-local message of the last failure */
const char*  zlb_version(void);

/* page-locked host buffers (faster H2D/D2H for zlb_encode_blocks / zlb_decode_blocks); plain malloc'd memory
 * is accepted everywhere as well */
void*        zlb_host_alloc(size_t bytes);
void         zlb_host_free(void* p);

/* upper bound of the compressed size of n input bytes (for sizing `out`) */
size_t       zlb_encode_bound(size_t n);

/* ---- encode ------------------------------------------------------------------------------------------ */
zlb_encoder* zlb_encoder_begin(zlb_ctx* ctx, int level);   /* level 0..4 (src/libzling_lz.cpp:129-135) */
void         zlb_encoder_end(zlb_encoder* enc);

/* Encode n consecutive stream bytes held in HOST memory.  n must be a multiple of ZLB_BLOCK_BYTES except for
 * the final call of a stream, and n <= max_blocks * ZLB_BLOCK_BYTES.  Writes the framed bytes of exactly these
 * blocks (sub-block records + one stop byte per block) to out[0..*out_len).  Includes H2D and D2H copies. */
int zlb_encode_blocks(zlb_encoder* enc, const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len);

/* Same, but `d_in` / `d_out` are DEVICE pointers on the context's GPU (inputs already resident in HBM, output
 * left in HBM); only the per-sub-block size table crosses PCIe.  d_in must be 16-byte aligned and readable for
 * 16 bytes past n. */
int zlb_encode_blocks_device(zlb_encoder* enc, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t out_cap, size_t* out_len);

/* ---- batch: many independent streams in one call -------------------------------------------------------
 * The reference encodes one stream per Encode() call (demo/zling.cpp:195-213); a 16 MiB block is one CTA here, so a single short
 * stream leaves most of the GPU idle.  zlb_encode_batch takes whole streams (each starts from the initial MTF tables and from
 * current_level = level, exactly like a fresh Encode() call), lays their blocks side by side and runs ONE pass of the pipeline
 * over all of them: sum of ceil(n / 16 MiB) over the streams must not exceed max_blocks.  out_len is set per stream; the bytes
 * of every stream equal what Encode() produces for it alone.  zlb_decode_batch is the inverse (each `in` = a whole framed
 * stream of at most max_blocks blocks in total): one decode chain per stream runs concurrently. */
typedef struct {
    const uint8_t* in; size_t n;          /* host: the stream (encode: raw bytes; decode: framed bytes) */
    uint8_t* out; size_t out_cap;         /* host: where the result goes */
    size_t out_len;                       /* set by the call */
} zlb_stream_io;
int zlb_encode_batch(zlb_ctx* ctx, int level, zlb_stream_io* streams, int nstreams);
int zlb_decode_batch(zlb_ctx* ctx, zlb_stream_io* streams, int nstreams);

/* Split form for ONE stream sharded over several GPUs (contiguous block ranges per GPU).  The parse of a range does
 * not depend on the MTF tables carried from the previous range, and on the carried level only for its first
 * sub-block (src/libzling.cpp:185,261-266), so: submit = H2D + parse launch (returns at once, assumes the carried level
 * is the requested one); then install the previous range's final state with zlb_encoder_set_state while the parse
 * runs; complete = MTF ranks + Huffman + framing + D2H (re-parses the first block only if the carried level differs).
 * zlb_encode_blocks(...) == submit + complete. */
int zlb_encode_submit(zlb_encoder* enc, const uint8_t* in, size_t n);
int zlb_encode_complete(zlb_encoder* enc, uint8_t* out, size_t out_cap, size_t* out_len);

/* carried state: 65536 bytes of MTF tables (context-major, rank -> byte) followed by int32 LE current level */
int zlb_encoder_get_state(zlb_encoder* enc, uint8_t* state /* ZLB_STATE_BYTES */);
int zlb_encoder_set_state(zlb_encoder* enc, const uint8_t* state);

/* ---- multi-GPU: ONE stream over several GPUs, one process per GPU ---------------------------------------
 * The reference has no counterpart (it is single-threaded); what these calls move between GPUs is exactly the state the
 * reference keeps outside its block loop: m_mtf[256] (src/libzling_lz.h:105, never reset) and current_level
 * (src/libzling.cpp:185,261-266).  NCCL (libnccl.so.2) is loaded at run time on first use.
 *   zlb_comm_get_unique_id   rank 0 makes the id; the caller distributes it to all ranks (any transport)
 *   zlb_comm_create          ncclCommInitRank on the context's GPU; collective over all ranks
 *   zlb_encode_stream_sharded  rank r owns a contiguous range of 16 MiB blocks of the stream (ranges in rank order;
 *        `in`/`n` = this rank's range, n a multiple of ZLB_BLOCK_BYTES except on the last non-empty rank, n = 0 allowed).
 *        Every rank parses its range at once; the 65 540-byte carried state travels GPU -> GPU in block order
 *        (ncclRecv from rank-1, ncclSend to rank+1, device buffers); MTF ranks + Huffman + framing of a range run when
 *        its carried state has arrived; ONE variable-length gather (sizes: ncclAllGather of u64; payloads: one group of
 *        ncclSend/ncclRecv) brings the framed bytes to rank 0, which copies the whole stream to `out` (host; may be
 *        NULL on the other ranks).  *out_len = total stream bytes on rank 0, this rank's framed bytes elsewhere.
 *        Collective: every rank of the communicator must call it.  The bytes equal zlb_encode_blocks on one GPU.
 *   zlb_encode_blocks_gathered  independent streams, one per rank (zlb_encode_blocks on every rank), whose framed outputs stay
 *        in device memory and reach rank 0 through the same single gather: `out` on rank 0 = the streams back to back in
 *        rank order, sizes[r] = bytes of rank r's stream
 *   zlb_gather_packed        the gather alone, for outputs already in device memory */
#define ZLB_COMM_ID_BYTES 128
typedef struct zlb_comm zlb_comm;
typedef struct {
    double   ms_wait_carry;   /* exchange stream: from the parse launch to the arrival of the previous range's state */
    double   ms_gather;       /* sizes + payload gather (+ D2H of the whole stream on rank 0) */
    uint64_t local_bytes;     /* framed bytes of this rank's range */
    uint64_t total_bytes;     /* framed bytes of the whole stream */
    uint64_t spec_reparsed_blocks; /* blocks re-parsed for the level feedback before the carried state arrived (ranks > 0) */
} zlb_shard_stats;
int       zlb_comm_get_unique_id(uint8_t* id /* ZLB_COMM_ID_BYTES */);
zlb_comm* zlb_comm_create(zlb_ctx* ctx, int rank, int world, const uint8_t* id /* ZLB_COMM_ID_BYTES */);
void      zlb_comm_destroy(zlb_comm* comm);
int       zlb_comm_get_stats(const zlb_comm* comm, zlb_shard_stats* out);
int zlb_encode_stream_sharded(zlb_encoder* enc, zlb_comm* comm, const uint8_t* in, size_t n, int in_on_device,
                              uint8_t* out, size_t out_cap, size_t* out_len);
int zlb_encode_blocks_gathered(zlb_encoder* enc, zlb_comm* comm, const uint8_t* in, size_t n, int in_on_device,
                               uint8_t* out, size_t out_cap, size_t* out_len, uint64_t* sizes /* [world], may be NULL */);
int zlb_gather_packed(zlb_comm* comm, const uint8_t* d_local, size_t n_local, uint8_t* out, size_t out_cap, size_t* out_len,
                      uint64_t* sizes /* [world], may be NULL */);

/* ---- decode ------------------------------------------------------------------------------------------ */
zlb_decoder* zlb_decoder_begin(zlb_ctx* ctx);
void         zlb_decoder_end(zlb_decoder* dec);

/* Decode whole framed blocks held in HOST memory: `in` must start at a block start (a 0x01 flag) and contain
 * only complete blocks (each ending with its 0x00 stop flag), at most max_blocks of them.  *consumed = bytes of
 * `in` used, *out_len = decoded bytes written. */
int zlb_decode_blocks(zlb_decoder* dec, const uint8_t* in, size_t n, size_t* consumed, uint8_t* out, size_t out_cap, size_t* out_len);

/* ---- instrumentation (bench.py, tests) ------------------------------------------------------------------ */
typedef struct {
    double   ms_total;        /* device time of the last encode/decode call, first to last kernel (CUDA events) */
    double   ms_parse;        /* zl_rolz_parse (dominant kernel) */
    double   ms_mtf;
    double   ms_huff_build;
    double   ms_pack;
    double   ms_h2d, ms_d2h;
    uint32_t launches;        /* kernels launched by the last call */
    uint32_t parse_launches;
    uint32_t reparsed_blocks; /* blocks parsed again because the level-feedback prediction was wrong */
    uint64_t tokens;          /* tokens produced by the last call */
    uint64_t subblocks;
    uint64_t slow_main, slow_lazy, window_hits;   /* unused (kept for ABI stability of the struct) */
    uint64_t windows;         /* parse: windows of 1022 positions executed, summed over blocks */
    uint64_t cyc_spec;        /* parse: SM cycles spent in SPEC (records against the frozen bucket state), summed over blocks */
    uint64_t cyc_resolve;     /* parse: SM cycles spent in the fixed-point ROUNDS, summed over blocks */
    uint64_t general_path;    /* unused */
    uint64_t cyc_total;       /* parse: SM cycles from kernel start to end, summed over blocks */
    uint64_t flagged;         /* unused */
    uint64_t rounds;          /* parse: fixed-point rounds executed, summed over windows and blocks */
    uint64_t cyc_final;       /* parse: SM cycles in FINALIZE (bucket writes, token emission), summed over blocks */
    uint64_t cyc_orbit;       /* parse: SM cycles of the rounds spent on the orbit of the entry position */
    uint64_t cyc_rank;        /* parse: ... on the sub-block roll-over scan */
    uint64_t cyc_decide;      /* parse: ... on re-deriving the decisions of the marked positions */
} zlb_stats;
int zlb_get_stats(const zlb_ctx* ctx, zlb_stats* out);

/* intermediates of the last zlb_encode_blocks* call, for parity tests against the oracle:
 *   tokens of block `blk`: u32 per token = sym | aux<<10 | byte<<22 | raw<<31 (sym: literals carry the MTF rank after the
 *   call, aux = match idx or literal context), and the sub-block table (encpos_end, rlen, olen, level, ntok) */
typedef struct { uint32_t tok_begin, tok_end, enc_begin, enc_end, rlen, level, olen, bits_lo; } zlb_subblock;
int zlb_debug_tokens(zlb_ctx* ctx, int blk, uint32_t* tok, size_t cap, size_t* ntok);
int zlb_debug_subblocks(zlb_ctx* ctx, int blk, zlb_subblock* sub, size_t cap, size_t* nsub);

/* run only the Huffman table kernels on caller-supplied frequency tables (nsym = 514/cap 15 or 32/cap 8) */
int zlb_debug_huff_tables(zlb_ctx* ctx, const uint32_t* freq, int ntables, int nsym, int cap, uint8_t* len_out, uint16_t* code_out);

#ifdef __cplusplus
}
#endif
#endif /* ZLB_H */
