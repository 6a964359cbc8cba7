/* pvw_b200.h -- C ABI of the B200-native PVW hot path (libpvw_b200.so).
 *
 * This is the drop-in boundary a Rust shim of gnosisguild/pvw-rs binds with `extern "C"`.
 * The reference has no FFI of its own: the boundary is its public Rust API (src/lib.rs:31-55).  Every entry point
 * below names the reference interface it replaces (paths relative to the reference repository).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; the library never takes ownership of caller memory.
 *  - host polynomial layout = the reference's: one polynomial is u64[L][ell] row-major, canonical residues in
 *    [0, q_j), NTT representation (fhe-math Poly.coefficients, src/params/parameters.rs:455-458).
 *    Matrices / vectors of polynomials are plain C arrays of such blocks.
 *  - every function returns 0 on success or a negative pvw_status that maps 1:1 onto a PvwError variant
 *    (src/errors.rs:11-70); pvw_last_error() returns the message.  No exception or abort crosses the boundary.
 *  - a context is bound to ONE CUDA device and ONE shard of parties [row0, row0+nrows) (rows of the global public
 *    key B).  One process per GPU; the one exchange of the path (c1 slices) is pvw_shard_* below (copy engines over NVLink), or a
 *    collective of the host layer's choice (NCCL) on the device pointer pvw_ct_c1_device_ptr() exposes.  A context is not
 *    thread-safe; guard it with a mutex (the Rust shim does).
 *  - `PVW_IO_DEVICE` in `flags` means the data pointers of that call are CUDA device pointers on the context's
 *    device (inputs already resident in HBM); otherwise they are host pointers and the call performs the copies.
 *    Device-pointer calls are asynchronous: they are queued on pvw_ctx_stream() and return at once, so the caller must
 *    order its own producers / consumers of those buffers against that stream (events, or pvw_ctx_synchronize()).
 *    Host-pointer calls return when the caller's buffers may be reused and every host output has been written: a call without
 *    host outputs (pvw_encrypt_batch) returns once its inputs have been copied, while its kernels are still queued -- the
 *    ciphertexts live in the device store, and every later call of the context is ordered after them.
 *  - there is no CPU fallback: without a CUDA device pvw_ctx_create fails with PVW_ERR_INTERNAL.
 */
#ifndef PVW_B200_H
#define PVW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvw_ctx pvw_ctx;

typedef enum {
  PVW_OK = 0,
  PVW_ERR_INVALID_PARAMETERS = -1, /* PvwError::InvalidParameters   errors.rs:14 */
  PVW_ERR_DIMENSION_MISMATCH = -2, /* PvwError::DimensionMismatch   errors.rs:56 */
  PVW_ERR_INDEX_OUT_OF_BOUNDS = -3,/* PvwError::IndexOutOfBounds    errors.rs:59 */
  PVW_ERR_ENCRYPTION = -4,         /* PvwError::EncryptionError     errors.rs:20 */
  PVW_ERR_DECRYPTION = -5,         /* PvwError::DecryptionError     errors.rs:23 */
  PVW_ERR_KEYGEN = -6,             /* PvwError::KeyGenerationError  errors.rs:26 */
  PVW_ERR_INTERNAL = -7,           /* PvwError::InternalError       errors.rs:68 (CUDA failures, missing device) */
  PVW_ERR_DESERIALIZATION = -8,    /* PvwError::DeserializationError errors.rs:35 (malformed wire bytes)          */
  PVW_ERR_INSUFFICIENT_DATA = -9   /* PvwError::InsufficientData    errors.rs:62 (wire buffer too short)          */
} pvw_status;

enum { PVW_IO_HOST = 0u, PVW_IO_DEVICE = 1u,
       /* pvw_encrypt_batch only: compute just c1 (dealers c1_lo..c1_hi) or just c2, so that a multi-GPU host layer can start the
          all-gather of the c1 slices while the (much larger) c2 product runs; m / e2 may be NULL with C1_ONLY, e1 with C2_ONLY */
       PVW_ENC_C1_ONLY = 2u, PVW_ENC_C2_ONLY = 4u,
       /* pvw_encrypt_batch on a connected shard exchange (pvw_shard_connect): queue pvw_shard_push_c1 for the slice right after its
          product, inside the call, so that one call (host buffers included) overlaps the peer copies with the c2 product */
       PVW_ENC_PUSH_C1 = 8u,
       /* element type of the small signed inputs.  Default: int64_t, the reference's `i64` coefficient type (SecretKey.secret_coeffs,
          secret_key.rs:14-18; Poly::from_coefficients(&[i64]), encryption.rs:148).  The values are tiny -- CBD(variance <= 16) secrets
          and randomness lie in [-32, 32], errors within the bounds of parameters.rs:110-114 -- so a host that feeds many dealers per
          call can pass them narrow and cut the host-to-device volume 8x / 2-4x:
            PVW_IN_SECRET_I8: r (pvw_encrypt_batch), sk (pvw_decrypt_batch, pvw_keygen_batch) are int8_t
            PVW_IN_ERROR_I32 / _I16: e1, e2 (pvw_encrypt_batch), e (pvw_keygen_batch) are int32_t / int16_t
          The pointers are `const void *` for that reason; without a flag they point to int64_t as before. */
       PVW_IN_SECRET_I8 = 0x100u, PVW_IN_ERROR_I32 = 0x200u, PVW_IN_ERROR_I16 = 0x400u };

/* Replaces PvwParametersBuilder::build (src/params/parameters.rs:117-195) + fhe-math Context::new.
 * Same validation: n > 0, k > 0, ell a power of two >= 8 (8, 16, 32 run on register-resident kernels, 64 .. 256 on generic ones;
 * above 256 is rejected: no reference test, example or parameter recommendation goes past 32), moduli prime, < 2^62, = 1 mod 2*ell, distinct,
 * error bounds > 0 (bounds >= 2^63 are not representable here; the reference type is BigInt but every
 * call site uses <= u32, parameters.rs:110-114). */
typedef struct {
  uint32_t n;              /* number of parties                                   */
  uint32_t k;              /* LWE dimension                                       */
  uint32_t ell;            /* redundancy parameter = ring degree                  */
  uint32_t L;              /* number of RNS moduli                                */
  const uint64_t *moduli;  /* [L]                                                 */
  const uint64_t *psi;     /* [L] primitive 2*ell-th roots, or NULL = derive like fhe-math's NttOperator
                              (ChaCha8 seed 0 search; recalled)   */
  float secret_variance;   /* default 0.5  (parameters.rs:166); only recorded     */
  uint64_t error_bound_1;  /* default 100  (parameters.rs:167)                    */
  uint64_t error_bound_2;  /* default 200  (parameters.rs:168)                    */
  uint32_t row0;           /* first party (row of B) held by this context         */
  uint32_t nrows;          /* number of rows held; 0 = all n                      */
  int32_t device;          /* CUDA device ordinal                                 */
} pvw_params_desc;

int pvw_ctx_create(pvw_ctx **out, const pvw_params_desc *desc);
void pvw_ctx_destroy(pvw_ctx *ctx);
/* message of the last failing call on this context (or of the last failing pvw_ctx_create when ctx == NULL) */
const char *pvw_last_error(const pvw_ctx *ctx);

/* PvwParameters accessors: delta() / delta_power_l_minus_1() / q_total() (parameters.rs:369-386) as little-endian
 * u64 words (`cap` = capacity of out; *nwords receives the count), the psi actually used, and
 * verify_correctness_condition() (parameters.rs:510-551; returns 1/0 in *ok). which: 0 = Q, 1 = delta, 2 = delta^(ell-1) */
int pvw_params_bigint(const pvw_ctx *ctx, int which, uint64_t *out, uint32_t cap, uint32_t *nwords);
int pvw_params_psi(const pvw_ctx *ctx, uint64_t *out /* [L] */);
int pvw_params_correctness_condition(const pvw_ctx *ctx, int *ok);

/* PvwCrs storage (src/params/crs.rs:12-17): A is k x k polynomials in NTT form, A[i][j] at ((i*k + j)*L*ell). */
int pvw_crs_upload(pvw_ctx *ctx, const uint64_t *A /* [k][k][L][ell] */, uint32_t flags);
int pvw_crs_download(pvw_ctx *ctx, uint64_t *A /* [k][k][L][ell] */);

/* Seeded CRS: PvwCrs::new_deterministic (crs.rs:45-67) and new_from_tag (crs.rs:74-90).  Host-side expansion (as in the
 * reference: ChaCha8 master -> 32-byte element seeds -> fhe-math Poly::random_from_seed = SHA-256 + ChaCha8 + Uniform(0..q_j);
 * restated from memory, see csrc/crsgen.hpp) followed by the upload.  A_out (may be NULL) receives the matrix [k][k][L][ell]. */
int pvw_crs_generate_deterministic(pvw_ctx *ctx, const uint8_t seed[32], uint64_t *A_out);
int pvw_crs_generate_from_tag(pvw_ctx *ctx, const char *tag, uint64_t *A_out);
/* the same expansion without a context or a device (pure host arithmetic), and the tag -> seed map of new_from_tag */
int pvw_crs_expand_seed(uint32_t k, uint32_t ell, uint32_t L, const uint64_t *moduli, const uint8_t seed[32], uint64_t *A_out);
int pvw_crs_tag_to_seed(const char *tag, uint8_t seed_out[32]);

/* GlobalPublicKey storage (src/keys/public_key.rs:43-54): add_public_key (:214-250) for rows [row, row+count) of B.
 * `row` is a GLOBAL party index and must lie inside this context's shard.  num_keys = max(index)+1 as in :245-247. */
int pvw_pk_upload_rows(pvw_ctx *ctx, uint32_t row, uint32_t count, const uint64_t *B /* [count][k][L][ell] */, uint32_t flags);
int pvw_pk_download_rows(pvw_ctx *ctx, uint32_t row, uint32_t count, uint64_t *B /* [count][k][L][ell] */);
/* GlobalPublicKey::is_full (public_key.rs:349-351) restricted to this shard: every local row index < num_keys */
int pvw_pk_num_keys(const pvw_ctx *ctx, uint32_t *num_keys);

/* Batched key generation: PublicKey::generate (public_key.rs:111-147) + PvwCrs::multiply_by_secret_key
 * (crs.rs:138-171) for `count` parties starting at global index `row`, with the error e supplied explicitly:
 *   b_p[c] = sum_j NTT(s_p[j]) (.) A[j][c] + NTT(e_p[c]).   Rows land in the device-resident B. */
int pvw_keygen_batch(pvw_ctx *ctx, uint32_t row, uint32_t count, const void *sk /* [count][k][ell] int64_t (or PVW_IN_SECRET_I8) */,
                     const void *e /* [count][k][ell] int64_t (or PVW_IN_ERROR_*) */, uint32_t flags);

/* PvwCrs::multiply_by_randomness (crs.rs:177-205): out[i] = sum_j A[i][j] (.) r[j] for D independent vectors.
 * r_hat and out are NTT-form polynomials in the host layout.  len != k is the caller's DimensionMismatch. */
int pvw_crs_multiply_by_randomness(pvw_ctx *ctx, uint32_t D, const uint64_t *r_hat /* [D][k][L][ell] */,
                                   uint64_t *out /* [D][k][L][ell] */);

/* Device-resident ciphertext store: `capacity` dealers; slot d holds PvwCiphertext{c1 (k polys), c2 (nrows polys)}
 * (src/crypto/encryption.rs:15-24).  Re-reserving discards the contents. */
int pvw_ct_reserve(pvw_ctx *ctx, uint32_t capacity);

/* encrypt (encryption.rs:105-214) / encrypt_party_shares (:221-245) / encrypt_all_party_shares (:253-286) for D
 * dealers with the randomness the reference draws from thread_rng() passed in explicitly:
 *   c1_d[i] = sum_j A[i][j] (.) NTT(r_d[j]) + NTT(e1_d[i])
 *   c2_d[p] = sum_j B[p][j] (.) NTT(r_d[j]) + NTT((m_d[p] as i64) * g) + NTT(e2_d[p])      for local rows p
 * Results go to store slots [slot0, slot0+D).  c1 is computed only for dealers [c1_lo, c1_hi) of the batch (use
 * 0, D on a single GPU; with row sharding each rank computes a slice and the host layer all-gathers the rest into
 * the store, see pvw_ct_c1_device_ptr).  e1 may be NULL when c1_lo == c1_hi.
 * Fails like the reference when the key is not full (:117) or the correctness condition is false (:124). */
int pvw_encrypt_batch(pvw_ctx *ctx, uint32_t slot0, uint32_t D, uint32_t c1_lo, uint32_t c1_hi,
                      const uint64_t *m /* [D][nrows] */, const void *r /* [D][k][ell] int64_t (or PVW_IN_SECRET_I8) */,
                      const void *e1 /* [D][k][ell] int64_t (or PVW_IN_ERROR_*) */, const void *e2 /* [D][nrows][ell] likewise */,
                      uint32_t flags);

/* PvwCiphertext download / upload in the reference layout (c1 [k][L][ell], c2 [nrows][L][ell]); NULL = skip. */
int pvw_ct_download(pvw_ctx *ctx, uint32_t slot, uint64_t *c1, uint64_t *c2);
int pvw_ct_upload(pvw_ctx *ctx, uint32_t slot, const uint64_t *c1, const uint64_t *c2);
/* device address of c1 of store slot `slot` (layout [slot][L][k][ell], contiguous over slots: *slot_stride u64
 * elements apart) so that the host layer can run collectives (NCCL all-gather over dealers) in place. */
int pvw_ct_c1_device_ptr(pvw_ctx *ctx, uint32_t slot, void **ptr, uint64_t *slot_stride);

/* decrypt_party_value (src/crypto/decryption.rs:249-278) / decrypt_party_shares (:281-325) for P local parties
 * x D stored ciphertexts:  out[p*D + d] = decode( sum_j NTT(s_p[j]) (.) c1_d[j] - c2_d[party_idx[p]] ).
 * dealer_slots == NULL means slots 0..D-1.  party_idx are GLOBAL indices inside the shard.
 * This is the subset form of examples/pvw_valid_dec.rs:198-210; the "exactly n ciphertexts" rule of
 * decrypt_party_shares (:295) is enforced by the host layer. */
int pvw_decrypt_batch(pvw_ctx *ctx, uint32_t D, const uint32_t *dealer_slots /* [D] host, or NULL */, uint32_t P,
                      const uint32_t *party_idx /* [P] host */, const void *sk /* [P][k][ell] int64_t (or PVW_IN_SECRET_I8) */,
                      uint64_t *out /* [P][D] */, uint32_t flags);

/* decode_scalar_pvw_rns (decryption.rs:10-58) on `count` noisy polynomials given in the host layout. */
int pvw_decode_batch(pvw_ctx *ctx, uint32_t count, const uint64_t *zhat /* [count][L][ell] */, uint64_t *out /* [count] */);

/* small signed coefficients -> NTT form: Poly::from_coefficients + change_representation(Ntt)
 * (encryption.rs:147-154, secret_key.rs:98-112, parameters.rs:264-284). */
int pvw_ntt_forward_small(pvw_ctx *ctx, uint32_t count, const int64_t *coeffs /* [count][ell] */,
                          uint64_t *out /* [count][L][ell] */);

/* PvwParameters::encode_scalar (parameters.rs:346-367) for `count` scalars: out = NTT((m as i64) * [1, D, .., D^(l-1)]) */
int pvw_encode_scalars(pvw_ctx *ctx, uint32_t count, const uint64_t *m /* [count] */, uint64_t *out /* [count][L][ell] */);

/* ---- wire format (the crate's `serde` feature; bincode 1.3 of the hand-written Serialize impls) -------------------
 * A polynomial travels as a bincode `Vec<u8>` holding fhe-math's `Poly::to_bytes` (protobuf `Rq`: representation = NTT,
 * degree, residues bit-packed at bit_length(q_j - 1) bits, limb after limb).  For one parameter set every such record
 * has the same size, so every struct below has a fixed layout, returned by pvw_wire_layout_get():
 *   PvwCiphertext    (src/crypto/encryption.rs:298-354)  [u64 k][k records] [u64 n][n records] [params]
 *   public-key row   (src/keys/public_key.rs:471-487, :522-537)  [u64 k][k records]  = PublicKey.key_polynomials = one
 *                    row of GlobalPublicKey.matrix
 *   PvwCrs           (src/params/crs.rs:228-249)         [u64 k] k x ([u64 k][k records]) [params]
 *   PvwParameters    (src/params/parameters.rs:606-623)  n, k, l as u64; Vec<u64> moduli; f32 variance; two decimal strings
 * The serialisers run on the device where the data lives; deserialisers validate (lengths, Rq header, embedded
 * parameters == this context's, residues < q_j) before they write anything and fail with PVW_ERR_DESERIALIZATION.
 * Only the canonical encoding the reference's own serialiser emits is accepted.  The third-party encodings (bincode,
 * prost, fhe-util transcode) are restated from their published behaviour: parity unpinned, see DESIGN.md.
 * With PVW_IO_DEVICE the byte buffers are device pointers and the calls are asynchronous, except that every
 * deserialiser synchronises once to read the validation verdict. */
typedef struct {
  uint64_t poly_bytes;        /* Poly::to_bytes length                                   */
  uint64_t record_bytes;      /* 8 + poly_bytes: one element of a Vec<Vec<u8>>           */
  uint64_t params_bytes;      /* bincode(PvwParameters)                                  */
  uint64_t pk_row_bytes;      /* 8 + k * record_bytes                                    */
  uint64_t ciphertext_bytes;  /* bincode(PvwCiphertext) for all n parties                */
  uint64_t crs_bytes;         /* bincode(PvwCrs)                                         */
  uint64_t ct_c1_offset, ct_c2_offset, ct_params_offset; /* sections of a ciphertext blob; the record of party p starts at
                                                            ct_c2_offset + 8 + p * record_bytes */
} pvw_wire_layout;
int pvw_wire_layout_get(const pvw_ctx *ctx, pvw_wire_layout *out);
int pvw_wire_params(const pvw_ctx *ctx, uint8_t *out, uint64_t cap);
/* bincode::serialize(&PvwCiphertext) (encryption.rs:298-317) for store slots [slot0, slot0 + D): blob d at out + d*stride.
 * A row-sharded context writes the envelope, c1 and the c2 records of its own parties only (host output: the other
 * records are zero-filled; device output: left untouched), so that the ranks' byte ranges can simply be concatenated. */
int pvw_wire_ct_serialize(pvw_ctx *ctx, uint32_t slot0, uint32_t D, uint8_t *out, uint64_t stride, uint32_t flags);
/* bincode::deserialize::<PvwCiphertext> (encryption.rs:319-354) into store slots; a sharded context reads c1 and its own
 * parties' c2 records. */
int pvw_wire_ct_deserialize(pvw_ctx *ctx, uint32_t slot0, uint32_t D, const uint8_t *in, uint64_t stride, uint32_t flags);
/* rows [row, row + count) of GlobalPublicKey.matrix as `count` consecutive Vec<Vec<u8>> (public_key.rs:528-533, :471-487) */
int pvw_wire_pk_serialize_rows(pvw_ctx *ctx, uint32_t row, uint32_t count, uint8_t *out, uint32_t flags);
int pvw_wire_pk_deserialize_rows(pvw_ctx *ctx, uint32_t row, uint32_t count, const uint8_t *in, uint32_t flags);
/* bincode::serialize(&PvwCrs) / deserialize (crs.rs:228-295) */
int pvw_wire_crs_serialize(pvw_ctx *ctx, uint8_t *out, uint64_t cap, uint32_t flags);
int pvw_wire_crs_deserialize(pvw_ctx *ctx, const uint8_t *in, uint64_t len, uint32_t flags);
/* `count` polynomials in the host layout <-> `count` records (e.g. GlobalPublicKey.error_polynomials); host pointers */
int pvw_wire_polys_serialize(pvw_ctx *ctx, uint32_t count, const uint64_t *polys /* [count][L][ell] */, uint8_t *out);
int pvw_wire_polys_deserialize(pvw_ctx *ctx, uint32_t count, const uint8_t *in, uint64_t *polys /* [count][L][ell] */);

/* ---- multi-GPU: the c1 exchange of the row-sharded contexts of one box (one process per GPU) --------------------
 * Rows of B shard across the GPUs; c1 = A r + e1 does not.  Every rank computes c1 for its slice of a step's dealers
 * (pvw_encrypt_batch c1_lo..c1_hi) and needs all of it to decrypt (decryption.rs:257-263 reads the whole c1).  The reference has
 * no counterpart (single address space).  The exchange below runs on the COPY ENGINES over NVLink, under the c2 product, instead of
 * an all-gather kernel that competes with it for the SMs:
 *   setup   every rank: pvw_ct_reserve, pvw_shard_export(world) -> one pvw_shard_handle (CUDA IPC handles of its c1 store and of a
 *           small flag array); the host layer all-gathers the `world` handles by any means (MPI, torch.distributed, a file) and gives
 *           every rank the whole table: pvw_shard_connect(world, rank, handles).
 *   step    pvw_encrypt_batch(... c1_lo, c1_hi ...); pvw_shard_push_c1(slot0 + c1_lo, c1_hi - c1_lo) queues, on a dedicated stream and
 *           after the c1 product, one peer copy of the slice into every peer's store and an 8-byte counter write per peer;
 *           pvw_shard_wait_c1() makes the context's compute stream wait (cuStreamWaitValue64) until every peer's slice of this
 *           exchange has landed; then pvw_decrypt_batch; pvw_shard_release_c1() tells the peers that this rank's store may be
 *           overwritten by the next exchange (their next push waits for it).  Every rank issues the same sequence of push / wait /
 *           release calls; nothing synchronises the host.  The c1 product must be queued BEFORE the push, the c2 product after it
 *           (PVW_ENC_C1_ONLY, then PVW_ENC_C2_ONLY) if the transfer is to overlap the c2 product.
 *   pvw_ct_reserve is refused while connected (peers hold mappings of the store): pvw_shard_disconnect on every rank first. */
typedef struct { uint8_t bytes[192]; } pvw_shard_handle;
int pvw_shard_export(pvw_ctx *ctx, uint32_t world, pvw_shard_handle *out);
int pvw_shard_connect(pvw_ctx *ctx, uint32_t world, uint32_t rank, const pvw_shard_handle *all /* [world] */);
int pvw_shard_push_c1(pvw_ctx *ctx, uint32_t slot0, uint32_t count);
int pvw_shard_wait_c1(pvw_ctx *ctx);
int pvw_shard_release_c1(pvw_ctx *ctx);
int pvw_shard_disconnect(pvw_ctx *ctx);

/* blocks until all work queued by this context has finished; returns a sticky CUDA error if one occurred */
int pvw_ctx_synchronize(pvw_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of the context is ordered on, for CUDA-event timing by the caller */
void *pvw_ctx_stream(pvw_ctx *ctx);
/* tuning / introspection: "imma" (1 = batched products on the INT8 tensor cores, the default; 0 = CUDA-core kernel),
 * "imma_min_dealers" / "imma_min_rows" (smallest batch of dealers / rows of the matrix operand that take the tensor-core path,
 * defaults 8 / 16), "imma_chunk_dealers" (dealers per scratch
 * chunk, default 512), "imma_pair" (1 = the two-SM cta_group::2 form, a measured alternative), "imma_stages" (depth of the
 * product's shared-memory ring, 0 = as deep as fits), "imma_epilogue_warps" (8 = default, 16), "imma_fast_reduce" (1 = default: the one-step
 * reduction of the product's 160-bit sums when every modulus is >= 2^61), "ternary_tables" (1 = default: at ring degrees 8 and 16, secrets /
 * randomness with coefficients in {-1, 0, 1} are transformed by table lookup instead of butterflies; same results), "narrow_inputs" (1 = default:
 * 64-bit secrets of a batched decryption are copied to one byte per coefficient on the device before the per-limb transforms, when they fit), "planes_only" (1 = once the byte planes of B exist, free the u64 operand copy of B --
 * one resident copy instead of two; it is rebuilt from the planes when a single call, a download or a key update needs it), "gemm_impl" (CUDA-core kernel:
 * 0 = synchronous tiles, 1 = TMA bulk-copy pipeline, 2 = tensor-map boxes), "gemm_tile", "refill_lag", "tail_impl", "lift_fast", "decode_fused" (1 = fused decode of clean shares
 * with the general chain as the per-share fallback, the default -- at ring degrees 8 and 16 as two launches, short lift + carry chain, then the
 * claim check; 0 = the general chain for every share),
 * "decrypt_chunk_shares", "upload_chunk_bytes", "profile" */
int pvw_ctx_set_option(pvw_ctx *ctx, const char *name, int64_t value);
/* per-kernel-kind device timing, measured with CUDA events on the context's stream around every launch while the
 * option "profile" is 1 (2 = enable and reset, 0 = disable and reset): total milliseconds, launches and algorithmic
 * bytes accumulated since the last reset.  Synchronises the stream. */
enum { PVW_KERNEL_NTT = 0, PVW_KERNEL_MAC = 1, PVW_KERNEL_DECODE_RNS = 2, PVW_KERNEL_CRT_LIFT = 3, PVW_KERNEL_DECODE_TAIL = 4,
       PVW_KERNEL_PERMUTE = 5, PVW_KERNEL_WIRE = 6, PVW_KERNEL_EXPAND = 7, PVW_KERNEL_DECODE_FUSED = 8, PVW_KERNEL_IMMA = 9 /* the product on the INT8 tensor cores; PVW_KERNEL_MAC = on the CUDA cores */,
       PVW_KERNEL_NTT_PLANES = 10 /* secrets / randomness -> byte planes (table lookup for ternary polynomials, else butterflies); PVW_KERNEL_NTT = the other transforms */,
       PVW_KERNEL_KINDS = 11 };
int pvw_ctx_profile(pvw_ctx *ctx, int kind, double *ms_total, uint64_t *launches, double *algorithmic_bytes);
/* number of kernels launched by this context so far */
uint64_t pvw_ctx_launch_count(const pvw_ctx *ctx);
const char *pvw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PVW_B200_H */
