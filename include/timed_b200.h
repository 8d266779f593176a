/* timed_b200.h -- C ABI of libtimed_b200.so (B200 / sm_100a only).
 *
 * Drop-in boundary for the one hot path of wells-wood-research/timed-design: per-residue
 * 3D-CNN inference (what `tf.keras.models.load_model(...)` + `frame_model.predict(X_batch)`
 * do at /root/reference/predict.py:121,142) and the Monte-Carlo sequence sampler
 * (/root/reference/design_utils/sampling_utils.py:53-90,139-161).  The reference is pure
 * Python and has no FFI of its own; these entry points are what a ctypes binding placed at
 * those two call sites would bind.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch/Python types cross this boundary;
 *   - every function returns 0 on success, a negative code on failure, and
 *     timed_b200_last_error() then returns a thread-local message;
 *   - `cuda_stream` is a cudaStream_t passed as void* (NULL = default stream); calls enqueue
 *     work and return without synchronising unless stated otherwise;
 *   - the caller owns every buffer it passes; the library owns only what *_create returns;
 *   - a handle is bound to one device and must be driven by one host thread at a time.
 */
#ifndef TIMED_B200_H
#define TIMED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIMED_B200_ABI_VERSION 4 /* 2: sample_chains, consensus, seq_metrics, op_kernel, host-side I/O helpers; 3: predict_stats, float16 frames, voxelise; 4: pdb_parse, inflate_device, hdf5_frame_index (all additive) */

/* error codes */
#define TB_OK 0
#define TB_ERR_INVALID (-1) /* bad argument / unsupported graph */
#define TB_ERR_CUDA (-2)    /* a CUDA runtime/driver call failed */
#define TB_ERR_NO_DEVICE (-3)

/* element types of host/device frame buffers handed to graph_forward */
#define TB_DTYPE_F32 0
#define TB_DTYPE_F64 1
#define TB_DTYPE_U8 2 /* numpy bool / uint8 voxels (voxels_as_gaussian == False) */
#define TB_DTYPE_F16 3 /* IEEE half frames: halves the host->device bytes of predict_host (the e2e path is PCIe-bound at
                        * 222 KB of float32 per frame); the caller decides whether its frames are exact in half */

/* activation codes */
#define TB_ACT_NONE 0
#define TB_ACT_RELU 1
#define TB_ACT_ELU 2
#define TB_ACT_SIGMOID 3
#define TB_ACT_TANH 4

/* fused-op kinds (one tb_op_desc per op; ids are indices into the op array) */
#define TB_OP_INPUT 0   /* the (n,D,H,W,C) frame tensor                                        */
#define TB_OP_CONV3D 1  /* Conv3D / Dense / Flatten+Dense: act2(scale*act1(conv(x)+bias)+shift) */
#define TB_OP_POOL3D 2  /* Max/AveragePooling3D (TF 'same'/'valid' semantics)                  */
#define TB_OP_AFFINE 3  /* standalone BatchNormalization and/or activation                     */
#define TB_OP_GPOOL 4   /* GlobalAverage/MaxPooling3D                                          */
#define TB_OP_SOFTMAX 5 /* softmax over channels                                               */
#define TB_OP_CONCAT 6  /* channel concatenation                                               */
#define TB_OP_ADD 7     /* elementwise sum                                                     */

#define TB_MAX_INPUTS 8

/* One fused op of the inference graph.  Mirrors the Keras layer semantics listed in SURVEY.md
 * App. D (Keras 2.13 defaults): Conv3D is cross-correlation NDHWC x DHWIO, stride 1,
 * 'same' pads before=(k-1)/2 / after=k-1-before; pooling 'same' pads at the end, max ignores
 * padding, average divides by the number of valid elements. */
typedef struct tb_op_desc {
    int32_t op;                    /* TB_OP_*                                                 */
    int32_t n_inputs;              /* number of valid entries of inputs[]                     */
    int32_t inputs[TB_MAX_INPUTS]; /* producer op ids                                         */
    int32_t kernel[3];             /* CONV3D: (kd,kh,kw);  POOL3D: pool size                  */
    int32_t stride[3];             /* POOL3D strides (CONV3D: must be 1,1,1)                  */
    int32_t pad_same;              /* 1 = 'same', 0 = 'valid'                                 */
    int32_t c_out;                 /* CONV3D: filters / Dense units                           */
    int32_t pool_kind;             /* POOL3D / GPOOL: 0 = max, 1 = average                    */
    int32_t act1;                  /* TB_ACT_* applied to conv(x)+bias (AFFINE: before scale) */
    int32_t act2;                  /* TB_ACT_* applied after scale/shift                      */
    float alpha1;                  /* ELU alpha of act1                                       */
    float alpha2;                  /* ELU alpha of act2                                       */
    /* host float32 arrays, NULL when absent; copied (and re-laid-out) by graph_create:        */
    const float* kernel_w; /* CONV3D: (kd,kh,kw,C_in,C_out) row-major (Keras DHWIO)            */
    const float* bias;     /* CONV3D: (C_out)                                                  */
    const float* scale;    /* CONV3D/AFFINE: per-channel multiplier (folded BN gamma/sqrt(var+eps)) */
    const float* shift;    /* CONV3D/AFFINE: per-channel offset (folded BN beta - mean*scale)   */
} tb_op_desc;

typedef struct tb_graph tb_graph;

/* ---- library -------------------------------------------------------------------------------- */
int timed_b200_abi_version(void);
const char* timed_b200_last_error(void);
/* number of CUDA devices visible (0 when none; never fails) */
int timed_b200_device_count(void);

/* ---- inference graph  (replaces tf.keras.models.load_model / Model.predict,
 *      /root/reference/predict.py:121 and :142) ------------------------------------------------ */
/* Build a graph on `device` from `n_ops` fused ops; ops[0] must be TB_OP_INPUT with
 * kernel = (D,H,W) and c_out = C; the last op is the output (n, classes).  Weights are
 * packed (bf16 hi/lo planes, K-major) and uploaded here. */
int timed_b200_graph_create(const tb_op_desc* ops, int32_t n_ops, int32_t device, tb_graph** out);
void timed_b200_graph_destroy(tb_graph* g);
/* output width (classes) and algorithmic FLOPs per frame (2*D*H*W*k^3*Cin*Cout summed over
 * conv/dense ops) */
int timed_b200_graph_info(const tb_graph* g, int32_t* n_classes, double* flops_per_frame,
                          int32_t* n_kernel_launches_per_forward);
/* Numerics option.  1 (default): every conv keeps the two correction products of the bf16 split in their own TMEM
 * accumulator (3x less accumulator truncation; max |dp| 2.2e-5 on the TIMED-20 stand-in); wide tiles then hold ONE
 * accumulator stage, which their epilogue drains into registers and releases before the activation math.
 * 0: corrections accumulate into the main accumulator (max |dp| 5.5e-5, error proportional to the network's logit
 * gain) -- kept as an A/B switch.  Takes effect on the next forward. */
int timed_b200_graph_set_precise(tb_graph* g, int32_t precise);
/* Name of the CUDA kernel(s) op `op` launches for a forward of `n_frames` frames (the tile configuration, and
 * with it the kernel, depends on the frame count); NUL-terminated into buf. */
int timed_b200_graph_op_kernel(const tb_graph* g, int32_t op, int64_t n_frames, char* buf, int32_t buflen);
/* number of fused ops in the graph */
int timed_b200_graph_op_count(const tb_graph* g, int32_t* n_ops);
/* Per-op device timing for bench.py's roofline line.  While enabled, every graph_forward records
 * a CUDA event on its stream before the first op and after each op (up to 512 forwards).
 * read_op_times waits for the recorded forwards, writes the SUM of elapsed milliseconds per op
 * to ms_per_op[n_ops] (plus op kinds and algorithmic FLOPs/frame per op when non-NULL), the
 * number of forwards summed to n_forwards, and clears the record. */
int timed_b200_graph_set_timing(tb_graph* g, int32_t enabled);
int timed_b200_graph_read_op_times(tb_graph* g, float* ms_per_op, int32_t* op_kinds,
                                   double* op_flops_per_frame, int32_t* n_forwards);
/* device workspace needed to run `n_frames` frames in one forward */
int timed_b200_graph_workspace_bytes(const tb_graph* g, int64_t n_frames, size_t* out);
/* Forward `n_frames` frames resident on the device.  d_frames: (n,D,H,W,C) of `frames_dtype`
 * (cast to float32 first, as Keras does); d_probs: (n, classes) float32.  No allocation, no
 * synchronisation: capturable in a CUDA graph. */
int timed_b200_graph_forward(tb_graph* g, const void* d_frames, int32_t frames_dtype,
                             int64_t n_frames, void* d_workspace, size_t workspace_bytes,
                             float* d_probs, void* cuda_stream);
/* Same through HOST buffers (the call Model.predict maps to): copies frames host->device in
 * chunks, runs forward, copies probabilities back, synchronises.  Library-owned staging. */
int timed_b200_graph_predict_host(tb_graph* g, const void* h_frames, int32_t frames_dtype,
                                  int64_t n_frames, float* h_probs, int64_t max_chunk_frames);
/* How the last predict_host call was chunked: number of device passes and the largest pass in frames (never above
 * max_chunk_frames: Keras' batch_size bound, /root/reference/predict.py:142 via Model.predict(batch_size=...)). */
int timed_b200_graph_predict_stats(const tb_graph* g, int64_t* n_passes, int64_t* max_pass_frames);

/* ---- voxeliser: structure -> residue frames (the step BEFORE the path; aposteriori's make-frame-dataset, un-vendored:
 *      parameters at /root/reference/README.md:84-97,242 and /root/reference/ui.py:63-87, parity unpinned) ------------
 * Frames of the residues d_res_index[res_first .. res_first + n_res) of one structure (NULL index = identity) into
 * d_frames (n_res, V, V, V, C).  Atom table:
 * (n_atoms, 4) float32 x, y, z, gaussian sigma in voxel units; per-atom channel (< 0: not encoded), residue index and
 * C-beta flag.  d_res_frame: (residues, 12) float32 = origin (C-alpha) then the rows of the rotation into the residue's
 * local frame.  With encode_cb the centre residue's own C-beta is replaced by the ideal one (x, y, z, sigma in local
 * coordinates).  d_res_atom_range (residues, 2), optional: [first, end) of the atoms a residue's frame looks at, so that
 * several structures can share one atom table and one launch.  property_channel >= 0 adds d_res_property[residue] at C-beta positions to that channel.  d_scratch:
 * n_res * V^3 * C int32 (fixed-point accumulation: the result is independent of the order atoms are added in). */
int timed_b200_voxelise(const float* d_atoms_xyzs, const int32_t* d_atom_channel, const int32_t* d_atom_residue,
                        const int32_t* d_atom_is_cb, int64_t n_atoms, const float* d_res_frame, const float* d_res_property,
                        const int32_t* d_res_index, const int32_t* d_res_atom_range, int64_t res_first, int64_t n_res,
                        int32_t voxels_per_side, float voxel_edge, int32_t n_channels,
                        int32_t as_gaussian, int32_t encode_cb, const float* ideal_cb_xyz_sigma, int32_t cb_channel,
                        int32_t property_channel, int32_t* d_scratch, void* d_frames, int32_t frames_dtype, void* cuda_stream);

/* ---- single-layer entry for unit/parity tests -------------------------------------------------
 * y = act2(scale*act1(conv3d(x)+bias)+shift) on device buffers.
 * x: (n,D,H,W,C_in) float32 device; y: (n,Do,Ho,Wo,C_out) float32 device. Synchronises. */
int timed_b200_conv3d_fwd(const float* d_x, int64_t n, int32_t D, int32_t H, int32_t W,
                          int32_t c_in, const tb_op_desc* conv, int32_t device, float* d_y);

/* ---- Monte-Carlo sampler  (replaces design_utils/sampling_utils.py:53-90,139-161) ------------ */
/* p ** (1/t) / rowsum  in float64 (apply_temp_to_probs, sampling_utils.py:159-161).
 * d_probs_in/out: (n_rows, n_cls) float64 device (may alias). */
int timed_b200_apply_temperature(const double* d_probs_in, int64_t n_rows, int32_t n_cls,
                                 double t, double* d_probs_out, void* cuda_stream);
/* sequential float64 cumsum along axis 1 (probs.cumsum(axis=1), sampling_utils.py:82) */
int timed_b200_cumsum_rows(const double* d_probs, int64_t n_rows, int32_t n_cls, double* d_cdf,
                           void* cuda_stream);
/* Draw n_samples sequences of n_res residues: idx = first j with cdf[res][j] > r, else 0
 * (the (cumsum > r).argmax quirk, sampling_utils.py:82).  r comes from d_uniforms
 * ((n_samples,n_res) float64, parity hook for np.random.rand values) when non-NULL, else from
 * Philox4x32-10 keyed (seed, stream_id) with counter (sample, residue): independent of launch
 * geometry and of how samples are sharded over GPUs (pass first_sample = global index).
 * d_cls_to_letter: n_cls ASCII codes; d_seqs: (n_samples,n_res) uint8; d_idx (optional, may be
 * NULL): (n_samples,n_res) int32. */
int timed_b200_sample(const double* d_cdf, int64_t n_res, int32_t n_cls, int64_t n_samples,
                      int64_t first_sample, uint64_t seed, uint64_t stream_id,
                      const double* d_uniforms, const uint8_t* d_cls_to_letter, uint8_t* d_seqs,
                      int32_t* d_idx, void* cuda_stream);
/* Every chain of a structure set in ONE launch (the reference fans chains out over a process pool,
 * sampling_utils.py:181-191).  d_cdf: CDFs of all chains concatenated, (row_off[n_chains], n_cls) float64; chain c owns rows
 * [d_row_off[c], d_row_off[c+1]) and the bytes [d_seq_off[c], d_seq_off[c] + n_samples*n_res_c) of d_seqs, laid out
 * (n_samples, n_res_c) like a per-chain call; d_seq_off has n_chains+1 ascending multiples of 4, the last one =
 * total_seq_bytes.  Chain c is keyed (seed, stream_id0 + c): byte-identical to timed_b200_sample with that stream id. */
int timed_b200_sample_chains(const double* d_cdf, const int64_t* d_row_off, const int64_t* d_seq_off, int32_t n_chains,
                             int64_t total_seq_bytes, int32_t n_cls, int64_t n_samples, int64_t first_sample,
                             uint64_t seed, uint64_t stream_id0, const uint8_t* d_cls_to_letter, uint8_t* d_seqs,
                             void* cuda_stream);
/* Philox uniforms exactly as timed_b200_sample would draw them (test hook). */
int timed_b200_sample_uniforms(int64_t n_res, int64_t n_samples, int64_t first_sample,
                               uint64_t seed, uint64_t stream_id, double* d_out,
                               void* cuda_stream);

/* ---- post-processing on device  (design_utils/utils.py:659,768 + predict.py:163) ------------- */
/* argmax of float16-rounded probabilities, first index wins ties.  d_probs (n, n_cls) float32;
 * d_idx (n) int32. */
int timed_b200_argmax_fp16(const float* d_probs, int64_t n, int32_t n_cls, int32_t* d_idx,
                           void* cuda_stream);

/* NMR consensus (design_utils/utils.py:694-713): the states of one structure are `n_states[g]` consecutive blocks of
 * `n_res[g]` rows starting at `first_row[g]`; output rows of group g start at `out_row0[g]` (ascending).  Consensus =
 * running pairwise mean in float16 arithmetic over the fp16-rounded probabilities, c <- fp16(fp16(c + p_s) / 2), exactly
 * what numpy does on the float16 matrix of predict.py:163; d_idx = its first-index argmax.
 * d_probs (n_rows, n_cls) float32; d_consensus_fp16 (n_out_rows, n_cls) IEEE half bits; d_idx (n_out_rows) int32. */
int timed_b200_consensus_fp16(const float* d_probs, int64_t n_rows, int32_t n_cls, const int64_t* d_first_row,
                              const int32_t* d_n_states, const int32_t* d_n_res, const int64_t* d_out_row0,
                              int32_t n_groups, int64_t n_out_rows, uint16_t* d_consensus_fp16, int32_t* d_idx,
                              void* cuda_stream);

/* ---- sequence metrics of sampled sequences  (design_utils/analyse_utils.py:351-371, called at
 *      sampling_utils.py:132) ------------------------------------------------------------------ */
/* One warp per sequence: 20-bin composition, then charge at the reference pH, isoelectric point (first grid pH with
 * minimal |charge|), molecular weight, molar extinction at 280 nm from host-built tables
 * [mw(20)|ext280(20)|q_ref(20)|term_ref|water|n_grid|grid(n)|term(grid)(n)|q(grid)(n x 20)] (float64, device).
 * d_seqs (n_seqs, n_res) ASCII; d_letter_lut: 256 entries, letter -> 0..19 or -1; d_out (n_seqs, 4) float64
 * (NaN row when a sequence holds a letter outside the table). */
int timed_b200_seq_metrics(const uint8_t* d_seqs, int64_t n_seqs, int64_t n_res, const int8_t* d_letter_lut,
                           const double* d_tables, int32_t n_table_doubles, double* d_out, void* cuda_stream);

/* ---- host-side frame I/O  (replaces the per-frame h5py reads of design_utils/utils.py:514-529) ------------------
 * Inflate (zlib) / unshuffle the chunks of many HDF5 frame datasets on `n_threads` host threads and scatter them,
 * cast to float32 (float64/float32/uint8 sources) or copied (uint8 -> uint8), into a dense (n_frames, *frame_dims)
 * host array.  No device work.  file_base: the memory-mapped file; chunk c lives at file_base + src_off[c]
 * (src_size[c] stored bytes), belongs to output frame dst_frame[c] and starts at element coordinates
 * origin[c*rank .. c*rank+rank) of that frame; chunk_dims / frame_dims have `rank` entries (edge chunks are clipped).
 * deflate: 1 when the chunks are gzip-compressed; shuffle_elem_size: 0, or the element size when the HDF5 shuffle
 * filter was applied before deflate. */
int timed_b200_inflate_chunks(const uint8_t* file_base, int64_t n_chunks, const int64_t* src_off, const int64_t* src_size,
                              const int64_t* dst_frame, const int32_t* origin, int32_t rank, const int32_t* chunk_dims,
                              const int32_t* frame_dims, int32_t deflate, int32_t shuffle_elem_size, int32_t src_dtype,
                              int32_t dst_dtype, void* dst, int32_t n_threads);

/* The same on the device, for datasets whose frames are stored as ONE deflate-filtered chunk each (no shuffle): the STORED
 * bytes are shipped (18 KB instead of 222 KB per frame on real structures) and every zlib stream is inflated by its own warp
 * (csrc/inflate.cuh) into d_out + s * out_bytes.  d_comp: the stored bytes; stream s occupies [d_off[s], d_off[s] + d_size[s]).
 * d_status[s]: 0 ok, else the stream is malformed / truncated / does not inflate to exactly out_bytes (the caller then takes
 * the host path).  All pointers are device pointers; asynchronous on cuda_stream. */
int timed_b200_inflate_device(const uint8_t* d_comp, int64_t n_streams, const int64_t* d_off, const int64_t* d_size,
                              int64_t out_bytes, void* d_out, int32_t* d_status, void* cuda_stream);

/* Index of n frame datasets of a memory-mapped HDF5 file in one call (replaces the per-frame object lookups of
 * design_utils/utils.py:514-529 on the device-inflate route): obj_addr[i] is the object-header address of frame i (from the
 * chain group's link table).  The caller has parsed ONE frame of the batch in full and passes its raw dataspace / datatype /
 * filter-pipeline messages, the header of its `encoded_residue` attribute message (everything before the attribute's data)
 * and the data length; a frame is accepted only when those bytes are identical, its layout is version-3 chunked with exactly
 * one unmasked chunk at the origin (one-leaf version-1 B-tree) and every read stays inside file_len.  Outputs: chunk_off /
 * chunk_size (absolute offset and stored size of the chunk), attr_data (n * attr_data_len bytes), status[i] (0 ok; else the
 * caller walks the batch with its own reader).  Version-1 object headers only.  Host threads, no device work. */
int timed_b200_hdf5_frame_index(const uint8_t* file_base, int64_t file_len, int64_t base_addr, int64_t n,
                                const int64_t* obj_addr, const uint8_t* tmpl_space, int32_t space_len,
                                const uint8_t* tmpl_type, int32_t type_len, const uint8_t* tmpl_filters, int32_t filters_len,
                                const uint8_t* tmpl_attr_hdr, int32_t attr_hdr_len, int32_t attr_data_len, int32_t rank,
                                int64_t* chunk_off, int64_t* chunk_size, uint8_t* attr_data, int32_t* status,
                                int32_t n_threads);

/* Text of a (rows, cols) float32/float64 host matrix exactly as numpy.savetxt(..., delimiter=",") prints it ("%.18e",
 * comma separated, one line per row: the rotamer dump of predict.py:145-146), formatted on `n_threads` host threads.
 * out must hold 26 bytes per number; *written receives the text length.  No device work. */
int timed_b200_format_csv_e18(const void* data, int32_t dtype, int64_t rows, int64_t cols, char* out, int64_t out_cap,
                              int64_t* written, int32_t n_threads);

/* Parse a comma-separated text matrix (the {model}.csv that /root/reference/sample.py:32-34 reads back with
 * np.genfromtxt) with strtod on `n_threads` host threads.  Call once with out == NULL to obtain rows / cols, then with a
 * (rows * cols) float64 buffer.  Fails with TB_ERR_INVALID when the text is not a rectangular matrix of numbers.
 * No device work. */
int timed_b200_parse_csv(const char* text, int64_t len, double* out, int64_t out_cap, int64_t* rows, int64_t* cols,
                         int32_t n_threads);

/* ---- host-side structure files  (the aposteriori front end of /root/reference/ui.py:73-86, /root/reference/README.md:84-97:
 * structure files -> residue frames) ---------------------------------------------------------------
 * Read `n_paths` PDB files (plain or .gz) on `n_threads` host threads into per-state atom / residue tables: ATOM records
 * only, states closed by ENDMDL (first state with atoms unless all_states), residues keyed by chain + resSeq + iCode in
 * order of first appearance, alternate locations resolved per residue, first occurrence of an atom name wins, residue
 * numbers repeated through insertion codes dropped (counted in state_dup).  No device work.
 * timed_b200_pdb_sizes: totals over all files.  timed_b200_pdb_export fills caller-allocated arrays:
 *   file_status[n_paths]        0 ok, 1 unreadable, 2 no ATOM record, 3 malformed coordinate field (such files have no state)
 *   state_file[n_states], state_res_off / state_atom_off[n_states + 1], state_dup[n_states]
 *   res_chain[n_res], res_id[n_res * 4], res_label[n_res * 3] (space padded), res_has_bb[n_res], res_bb[n_res * 9] (N, CA, C)
 *   atom_xyz[n_atoms * 3], atom_name[n_atoms] (0 N, 1 CA, 2 C, 3 O, 4 OXT, 5 CB), atom_res[n_atoms] (index within its state)
 * (only the backbone + C-beta atoms the frame encoders use are exported, in file order). */
typedef struct tb_pdb_batch tb_pdb_batch;
int timed_b200_pdb_parse(const char* const* paths, int32_t n_paths, int32_t all_states, int32_t n_threads, tb_pdb_batch** out);
int timed_b200_pdb_sizes(const tb_pdb_batch* batch, int64_t* n_states, int64_t* n_res, int64_t* n_atoms);
int timed_b200_pdb_export(const tb_pdb_batch* batch, int32_t* file_status, int32_t* state_file, int64_t* state_res_off,
                          int64_t* state_atom_off, int32_t* state_dup, char* res_chain, char* res_id, char* res_label,
                          uint8_t* res_has_bb, double* res_bb, double* atom_xyz, int32_t* atom_name, int32_t* atom_res);
void timed_b200_pdb_free(tb_pdb_batch* batch);

#ifdef __cplusplus
}
#endif
#endif /* TIMED_B200_H */
