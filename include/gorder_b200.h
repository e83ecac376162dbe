/*
 * gorder_b200.h — C ABI of the B200-native per-frame order-parameter engine.
 *
 * This is the drop-in boundary for gorder's hot path.  The reference has
 * no FFI today; the seam this ABI replaces is the generic frame loop
 *
 *     system.traj_iter_map_reduce::<Reader, SystemTopology, AnalysisError>(
 *         file, n_threads, analyze_frame, topology, Some("Master"), begin, end, step, ..)
 *                                                   reference: src/analysis/common.rs:283-339
 *     fn analyze_frame(frame: &System, data: &mut SystemTopology) -> Result<(), AnalysisError>
 *                                                   reference: src/analysis/common.rs:201-235
 *     impl ParallelTrajData for SystemTopology { reduce, initialize }
 *                                                   reference: src/analysis/topology/mod.rs:256-278
 *
 * Host side (Rust, unchanged) classifies molecules once, then
 *   gorder_gpu_create()   replaces SystemTopology::new + per-thread clone   (topology/mod.rs:70-118)
 *   gorder_gpu_submit()   replaces analyze_frame() for a batch of frames    (common.rs:201-235)
 *   gorder_gpu_finish()   replaces ParallelTrajData::reduce + the Add chain (topology/mod.rs:236-272)
 *   gorder_gpu_destroy()  drops the state
 * and back-fills AnalysisOrder / Map / AssignedLeaflets / NormalsStorage from GorderResults
 * before calling the untouched converter (presentation/converter.rs:52).
 *
 * Plain C: pointers + sizes only, no torch / CUDA types in the signatures.  All arrays passed
 * to gorder_gpu_create() are copied; the caller may free them when the call returns.
 *
 * Index convention: every atom index in this file is a "slot" = position of the atom inside the
 * coordinate frame handed to gorder_gpu_submit (the reference's Master group, common.rs:92-103,
 * renumbered densely by the host, or simply the absolute atom index when whole frames are passed).
 * Molecules of one molecule type are congruent (topology/molecule.rs:224-244), so an atom of
 * molecule m is addressed as  mol_base[m] + relative_index  (bond.rs:87-99).
 */
#ifndef GORDER_B200_H
#define GORDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GORDER_ABI_VERSION 2

/* ---- enums (int32 in the structs) ------------------------------------------------------- */

/* AnalysisType, reference: src/input/analysis.rs (AAOrder / CGOrder / UAOrder). AA and CG share
 * the bond engine (topology/bond.rs); UA uses the virtual-hydrogen engine (uaorder.rs). */
enum { GORDER_KIND_AA = 0, GORDER_KIND_CG = 1, GORDER_KIND_UA = 2 };

/* Axis, reference: src/input/axis.rs */
enum { GORDER_AXIS_X = 0, GORDER_AXIS_Y = 1, GORDER_AXIS_Z = 2 };

/* MembraneNormal, reference: src/input/membrane_normal.rs, analysis/normal.rs:31-35 */
enum { GORDER_NORMAL_STATIC = 0, GORDER_NORMAL_DYNAMIC = 1, GORDER_NORMAL_MANUAL = 2 };

/* LeafletClassification, reference: src/analysis/leaflets.rs:208-217. MANUAL covers FromFile /
 * FromMap / FromNdx / Clustering: the host computes the table. */
enum {
    GORDER_LEAFLET_NONE = 0,
    GORDER_LEAFLET_GLOBAL = 1,
    GORDER_LEAFLET_LOCAL = 2,
    GORDER_LEAFLET_INDIVIDUAL = 3,
    GORDER_LEAFLET_MANUAL = 4,
    /* spherical clustering (spherical_clustering.rs:36-275): 1-D two-component Gaussian mixture over the heads' distances
     * from the vesicle centre, outer cluster = upper.  `membrane` holds the ClusterHeads group; every analysed molecule's
     * head must be in it. */
    GORDER_LEAFLET_SPHERICAL = 5
};

/* Frequency, reference: src/input/frequency.rs; leaflets.rs:435-441 */
enum { GORDER_FREQ_EVERY = 0, GORDER_FREQ_ONCE = 1 };

/* Geometry, reference: src/analysis/geometry.rs:288,372,456 */
enum { GORDER_GEOM_NONE = 0, GORDER_GEOM_CUBOID = 1, GORDER_GEOM_CYLINDER = 2, GORDER_GEOM_SPHERE = 3 };

/* GeomReference, reference: src/input/geometry.rs (Point / Selection / Center) */
enum { GORDER_GEOMREF_POINT = 0, GORDER_GEOMREF_SELECTION = 1, GORDER_GEOMREF_BOX_CENTER = 2 };

/* Plane, reference: src/input/ordermap.rs:44-50. NOTE YZ projects to (z, y). */
enum { GORDER_PLANE_XY = 0, GORDER_PLANE_XZ = 1, GORDER_PLANE_YZ = 2 };

/* united-atom carbon kinds, reference: src/analysis/uaorder.rs:234-239 */
enum { GORDER_UA_CH3 = 0, GORDER_UA_CH2 = 1, GORDER_UA_CH1_UNSAT = 2, GORDER_UA_CH1_SAT = 3 };

/* Leaflet value in every table of this ABI: the reference's export convention (lib.rs:416-422). */
enum { GORDER_LOWER = 0, GORDER_UPPER = 1 };

/* Index of the three accumulators kept per order slot (bond.rs:221-247). */
enum { GORDER_TOTAL = 0, GORDER_ACC_UPPER = 1, GORDER_ACC_LOWER = 2 };

/* Error codes; 1:1 with AnalysisError (reference: src/errors.rs:121-169) + device errors. */
enum {
    GORDER_OK = 0,
    GORDER_ERR_UNDEFINED_BOX = 1,              /* AnalysisError::UndefinedBox */
    GORDER_ERR_NOT_ORTHOGONAL_BOX = 2,         /* AnalysisError::NotOrthogonalBox */
    GORDER_ERR_ZERO_BOX = 3,                   /* AnalysisError::ZeroBox */
    GORDER_ERR_UNDEFINED_POSITION = 4,         /* AnalysisError::UndefinedPosition(idx): NaN coordinate */
    GORDER_ERR_INVALID_GLOBAL_CENTER = 5,      /* AnalysisError::InvalidGlobalMembraneCenter */
    GORDER_ERR_INVALID_LOCAL_CENTER = 6,       /* AnalysisError::InvalidLocalMembraneCenter(head) */
    GORDER_ERR_MANUAL_LEAFLET_FRAME = 7,       /* ManualLeafletClassificationError::FrameNotFound */
    GORDER_ERR_DYNAMIC_NORMAL_POINTS = 8,      /* DynamicNormalError::NotEnoughPoints(n) */
    GORDER_ERR_DYNAMIC_NORMAL_SVD = 9,         /* DynamicNormalError::SVDFailed */
    GORDER_ERR_MANUAL_NORMAL_FRAME = 10,       /* ManualNormalError::FrameNotFound */
    GORDER_ERR_LEAFLET_FRAME_UNAVAILABLE = 11, /* shard does not hold the assignment frame (leaflets.rs:1529) */
    GORDER_ERR_ORDER_OVERFLOW = 12,            /* OrderValue overflowed (order.rs:47-60 panic) */
    GORDER_ERR_INVALID_ARGUMENT = 20,
    GORDER_ERR_ORDERMAP_BIN_TOO_LARGE = 21,    /* OrderMapConfigError::BinTooLarge */
    GORDER_ERR_ORDERMAP_NO_BOX = 22,           /* OrderMapConfigError::InvalidBoxAuto */
    GORDER_ERR_NO_DEVICE = 30,                 /* no usable CUDA device: there is NO CPU fallback */
    GORDER_ERR_CUDA = 31,
    GORDER_ERR_OUT_OF_MEMORY = 32,
    GORDER_ERR_NCCL = 33,                      /* NCCL missing at run time, or a collective failed */
    /* structure / topology / classification (host only; gorder_system_*, gorder_classify_*) */
    GORDER_ERR_IO = 40,                        /* file missing or unreadable (BondsError::FileNotFound / CouldNotReadLine) */
    GORDER_ERR_TPR_FORMAT = 41,                /* not a TPR file, unsupported tpx version, truncated or inconsistent */
    GORDER_ERR_BONDS_PARSE = 42,               /* BondsError::CouldNotParse */
    GORDER_ERR_BONDS_ATOM_NOT_FOUND = 43,      /* BondsError::AtomNotFound */
    GORDER_ERR_BONDS_SELF = 44,                /* BondsError::SelfBonding */
    GORDER_ERR_TOPOLOGY_NO_HEAD = 45,          /* TopologyError::NoHead(first atom of the molecule) */
    GORDER_ERR_TOPOLOGY_MULTIPLE_HEADS = 46,   /* TopologyError::MultipleHeads */
    GORDER_ERR_TOPOLOGY_NO_METHYL = 47,        /* TopologyError::NoMethyl */
    GORDER_ERR_TOPOLOGY_INCONSISTENT_METHYLS = 48, /* TopologyError::InconsistentNumberOfMethyls */
    GORDER_ERR_TOPOLOGY_NO_UA_CARBONS = 49,    /* TopologyError::NoUACarbons */
    GORDER_ERR_NO_TOPOLOGY = 50,               /* ConfigError::NoTopology: structure without bonds and no bonds file */
    GORDER_ERR_PDB_TOPOLOGY = 51,              /* ConfigError::InvalidPdbTopology: repeated atom numbers make CONECT ambiguous */
    GORDER_ERR_STRUCTURE_FORMAT = 52,          /* ConfigError::InvalidStructureFormat, or a GRO / PDB file that does not parse */
    GORDER_ERR_NDX_PARSE = 53,                 /* NdxLeafletClassificationError::CouldNotParse (groan ParseNdxError) */
    GORDER_ERR_NDX_INVALID_NAME = 54,          /* NdxLeafletClassificationError::InvalidName */
    GORDER_ERR_NDX_DUPLICATE_NAME = 55,        /* NdxLeafletClassificationError::DuplicateName */
    GORDER_ERR_NDX_GROUP_NOT_FOUND = 56,       /* NdxLeafletClassificationError::GroupNotFound */
    GORDER_ERR_NDX_ASSIGNMENT_NOT_FOUND = 57   /* NdxLeafletClassificationError::AssignmentNotFound(molecule, head) */
};

/* ---- setup ------------------------------------------------------------------------------ */

/* One molecule type (reference: topology/molecule.rs:147-169 MoleculeType<O>). */
typedef struct GorderMolType {
    int32_t n_molecules;
    const int32_t *mol_base;   /* [n_molecules] slot of relative index 0 of every molecule */

    /* AA / CG: bond types in the reference's sorted order (bond.rs:77-81); rel1 < rel2 so that
     * the vector goes from the lower to the higher absolute index (bond.rs:298-302). */
    int32_t n_bond_types;
    const int32_t *bond_rel;   /* [n_bond_types][2] */

    /* UA: carbon types sorted by relative index (topology/uatom.rs:39-41).
     * ua_rel[i] = {target, helper1, helper2, helper3}; helper3 = -1 unless CH1_SAT.
     * Roles as in uaorder.rs:911-915 (AtomTriplet) and :1050-1056 (AtomQuadruplet). */
    int32_t n_ua_atoms;
    const int32_t *ua_kind;    /* [n_ua_atoms] GORDER_UA_* */
    const int32_t *ua_rel;     /* [n_ua_atoms][4] */

    /* leaflets: head identifier (leaflets.rs:575,633,740) and methyls (leaflets.rs:743), relative */
    int32_t head_rel;          /* -1 if unused */
    int32_t n_methyls;
    const int32_t *methyl_rel; /* [n_methyls] */

    /* dynamic normals: this molecule's reference head (normal.rs:136), relative; -1 if unused */
    int32_t normal_head_rel;

    /* GORDER_LEAFLET_MANUAL: [n_manual_leaflet_frames][n_molecules], GORDER_UPPER / GORDER_LOWER,
     * BEFORE flip (leaflets.rs:816-874). Row = frame_index / real_frequency (0 for Once). */
    int32_t n_manual_leaflet_frames;
    const uint8_t *manual_leaflets;

    /* GORDER_NORMAL_MANUAL: [n_manual_normal_frames][n_molecules][3] (normal.rs:259-298).
     * Row = frame_index / step. */
    int32_t n_manual_normal_frames;
    const float *manual_normals;
} GorderMolType;

typedef struct GorderSetup {
    int32_t abi_version;       /* GORDER_ABI_VERSION */
    int32_t kind;              /* GORDER_KIND_* */
    int32_t n_atoms;           /* slots per frame */
    int32_t handle_pbc;        /* Analysis::handle_pbc (common.rs:207): 1 = PBC3D, 0 = NoPBC */
    int32_t step;              /* Analysis::step; frame_index passed to submit is a multiple of it */

    int32_t n_moltypes;
    const GorderMolType *moltypes;

    /* membrane normal (normal.rs:31-72) */
    int32_t normal_mode;       /* GORDER_NORMAL_* */
    int32_t normal_axis;       /* STATIC: GORDER_AXIS_* */
    float dynamic_radius;      /* DYNAMIC: radius of the head cloud, nm (normal.rs:140) */
    int32_t n_normal_heads;    /* DYNAMIC: group "NormalHeads" (common.rs:173-183): all atoms of it */
    const int32_t *normal_heads;
    int32_t collect_normals;   /* DYNAMIC: keep per-frame normals for export (normal.rs:211-227) */

    /* leaflets (leaflets.rs:208-330) */
    int32_t leaflet_mode;      /* GORDER_LEAFLET_* */
    int32_t leaflet_axis;      /* membrane normal used by the classifier (leaflets.rs:241-255) */
    int32_t leaflet_freq_kind; /* GORDER_FREQ_* */
    int32_t leaflet_freq;      /* EVERY: real frequency = configured frequency x step (leaflets.rs:261-262) */
    int32_t leaflet_flip;      /* leaflets.rs:68-73 */
    float leaflet_radius;      /* LOCAL: cylinder radius, nm */
    int32_t n_membrane;        /* GLOBAL / LOCAL: group "Membrane" */
    const int32_t *membrane;
    int32_t collect_leaflets;  /* keep assignment tables of every assignment frame for export */

    /* geometry selection (geometry.rs) */
    int32_t geom_kind;         /* GORDER_GEOM_* */
    int32_t geom_invert;
    int32_t geom_ref_kind;     /* GORDER_GEOMREF_* */
    float geom_ref_point[3];   /* POINT */
    int32_t n_geom_ref;        /* SELECTION: group "GeomReference" */
    const int32_t *geom_ref;
    float geom_dims[6];        /* CUBOID: xmin,xmax,ymin,ymax,zmin,zmax (+-INFINITY allowed);
                                  CYLINDER: radius, span_min, span_max;  SPHERE: radius */
    int32_t geom_axis;         /* CYLINDER orientation */
    float structure_box[3];    /* POINT reference: box of the STRUCTURE file.  The reference builds the shape of a fixed
                                  reference point ONCE, in GeometrySelection::new (geometry.rs:297-312), with the box of the
                                  structure file -- init_reference (geometry.rs:192-210) never rebuilds it -- so the wrap of
                                  the shape's origin uses that box while `inside` uses the frame's.  All zero: the frame's box
                                  is used (identical whenever reference + lower offsets lies inside the box). */

    /* order maps (ordermap.rs:40-96). Spans are resolved by the host exactly as Map::new does
     * (auto span = box of the structure file). */
    int32_t map_enabled;
    int32_t map_plane;         /* GORDER_PLANE_* */
    float map_span_x[2];       /* span of the first projected coordinate */
    float map_span_y[2];
    float map_bin[2];

    /* error estimation / convergence: keep per-frame sums (order.rs:83-88, timewise.rs:131-140) */
    int32_t timewise;

    /* engine knobs (no reference counterpart) */
    int32_t device;            /* CUDA device ordinal */
    int32_t max_batch_frames;  /* frames per launch; 0 = default */
} GorderSetup;

/* ---- results ---------------------------------------------------------------------------- */

/* Order slots: one per bond type (AA/CG) or per virtual C-H bond (UA: CH3 -> 3 slots, CH2 -> 2,
 * CH1 -> 1; uaorder.rs:253-272), concatenated over molecule types in input order.
 * All integer sums are the reference's fixed-point OrderValue (order.rs:13-26): round(S * 1e6). */
typedef struct GorderResults {
    int64_t n_slots;           /* out: total order slots */
    int64_t n_frames;          /* out: SystemTopology::total_frames (topology/mod.rs:49) */
    int64_t n_map_bins;        /* out: nx * ny (0 if maps disabled) */
    int64_t map_nx, map_ny;    /* out */
    int64_t n_leaflet_frames;  /* out: number of assignment frames held (collect_leaflets) */
    int64_t n_molecules_total; /* out: sum of n_molecules over molecule types */

    /* caller-allocated (may be NULL to skip); filled by gorder_gpu_finish / gorder_gpu_results */
    int64_t *sum;              /* [n_slots][3]   AnalysisOrder::order      (order.rs:73) */
    uint64_t *count;           /* [n_slots][3]   AnalysisOrder::n_samples  (order.rs:76) */
    int64_t *tw_sum;           /* [n_frames][n_slots][3]  TimeWiseData::order     (timewise.rs:134) */
    uint64_t *tw_count;        /* [n_frames][n_slots][3]  TimeWiseData::n_samples (timewise.rs:136) */
    int64_t *tw_frame_index;   /* [n_frames] frame_index of each row (ascending) */
    int64_t *map_sum;          /* [n_slots][3][n_map_bins]  Map::values  (ordermap.rs:27), x-major */
    uint64_t *map_count;       /* [n_slots][3][n_map_bins]  Map::samples (ordermap.rs:30) */
    uint8_t *leaflets;         /* [n_leaflet_frames][n_molecules_total] GORDER_UPPER/LOWER after flip */
    int64_t *leaflet_frame_index; /* [n_leaflet_frames] */
    float *normals;            /* [n_frames][n_molecules_total][3], NaN where not computed (normal.rs:217) */
} GorderResults;

typedef struct GorderHandle GorderHandle;

/* ---- entry points ----------------------------------------------------------------------- */

/* Build device state from the classified topology. Fails with GORDER_ERR_NO_DEVICE when no CUDA
 * device is usable: there is no CPU fallback in this library. */
int gorder_gpu_create(const GorderSetup *setup, GorderHandle **out);

/* Analyse n_frames frames (replaces n_frames calls of analyze_frame, common.rs:201-235).
 *   xyz         [n_frames][n_atoms][3] f32, nm, host memory (pinned memory makes the copy async);
 *               NaN marks an undefined position (-> GORDER_ERR_UNDEFINED_POSITION)
 *   box         [n_frames][3] orthogonal box lengths, nm; ignored if !handle_pbc.
 *               box validity (check_box, common.rs:186-198) is the host's job; an all-zero box
 *               is still rejected here with GORDER_ERR_ZERO_BOX.
 *   frame_index [n_frames] SystemTopology::frame of each frame (topology/mod.rs:41-43), i.e.
 *               (ordinal of the trajectory frame since `begin`), a multiple of `step`, strictly
 *               increasing across calls.
 * Returns when the batch has been queued and the input buffers may be reused. Errors detected on
 * the device are reported by a later submit or by finish (first error wins, as in the reference).
 * Thread safety: every entry point that takes a handle locks it, so several host threads (the reference's decode
 * threads) may share one handle; because frame_index must increase, they hand their batches over in trajectory order
 * (or each thread drives its own handle over its own frame range and the handles are combined like GPUs, see
 * gorder_gpu_accumulator_block). */
int gorder_gpu_submit(GorderHandle *h, const float *xyz, const float *box,
                      const int64_t *frame_index, int32_t n_frames);

/* Same, with frames already resident in device memory in the same [n_frames][n_atoms][3] layout. */
int gorder_gpu_submit_device(GorderHandle *h, const float *d_xyz, const float *d_box,
                             const int64_t *frame_index, int32_t n_frames);

/* Device-native frame layout ("planes"): frame_floats floats per frame; atom slot s,
 * component c lives at plane_offset[s] + c * plane_cstride[s]. A host that fills its pinned
 * buffers through this map (the Rust shim's per-frame copy of the Master group is a gather
 * anyway) can skip the device-side re-layout pass. */
int gorder_gpu_native_layout(GorderHandle *h, int64_t *frame_floats,
                             int32_t *plane_offset /* [n_atoms] or NULL */,
                             int32_t *plane_cstride /* [n_atoms] or NULL */);
int gorder_gpu_submit_native(GorderHandle *h, const float *planes_host, const float *box,
                             const int64_t *frame_index, int32_t n_frames);
int gorder_gpu_submit_native_device(GorderHandle *h, const float *d_planes, const float *d_box,
                                    const int64_t *frame_index, int32_t n_frames);

/* Optional: announce how many frames will be analysed in total (the reference knows the trajectory
 * length only at the end; with error estimation the per-frame rows, order.rs:83-88, then grow without
 * re-allocation). */
int gorder_gpu_reserve_frames(GorderHandle *h, int64_t n_frames);

/* Seed the leaflet assignment used until the next assignment frame (Frequency::Once on a shard
 * that does not own frame 0: SURVEY.md §8e). table: [n_molecules_total] after flip. */
int gorder_gpu_set_leaflets(GorderHandle *h, const uint8_t *table, int64_t frame_index);

/* Wait for all queued work, report the first deferred error. */
int gorder_gpu_sync(GorderHandle *h);

/* Sizes needed to allocate GorderResults arrays (fills the scalar "out" fields only). */
int gorder_gpu_result_sizes(GorderHandle *h, GorderResults *r);

/* Blocks; copies accumulators into the caller's arrays (replaces ParallelTrajData::reduce for the
 * single-GPU case). May be called more than once.  It does NOT reduce across GPUs: with frames sharded over several
 * handles, merge them first (gorder_gpu_reduce / gorder_gpu_reduce_comm below), then call finish on the root. */
int gorder_gpu_finish(GorderHandle *h, GorderResults *r);

/* ---- multi-GPU merge (replaces ParallelTrajData::reduce over the per-thread clones, topology/mod.rs:256-272, and the Add
 * chain bond.rs:449-465, order.rs:160-176, timewise.rs:34-51, ordermap.rs:116-138, normal.rs:234-256) ------
 * The analysed frames are sharded over GPUs in CONTIGUOUS ranges (range starts on multiples of the leaflet-assignment
 * period), one handle per GPU, all created from the same setup.  The merge leaves on the root handle
 *   - the SUM of the accumulator blocks (integer: exact, order-free),
 *   - the per-frame rows, collected leaflet tables and normals of all shards in frame order (shards ordered by the
 *     frame_index of their first frame),
 * so that gorder_gpu_finish(root) returns what a single handle fed with all frames would.  The first error of any shard (in
 * shard order) is returned, as the first Err aborts the reference's map-reduce.  The merge ends the analysis: do not submit
 * further frames to the root.
 *
 * One process drives all devices: one kernel on the root device reads the peers' blocks through NVLink peer access. */
int gorder_gpu_reduce(GorderHandle **handles, int32_t n, int32_t root);

/* One process per GPU (torchrun, MPI): NCCL, loaded at run time (libnccl.so.2; GORDER_NCCL_LIB overrides the name), on a
 * communicator the library creates: rank 0 calls gorder_comm_unique_id, the host broadcasts the 128 bytes by any means,
 * every rank calls gorder_comm_create.  gorder_gpu_reduce_comm is collective: ONE ncclReduce (int64 sum) of the block plus
 * grouped send / recv of the per-frame data to the root. */
#define GORDER_COMM_ID_BYTES 128
typedef struct GorderComm GorderComm;
int gorder_comm_unique_id(uint8_t *id /* [GORDER_COMM_ID_BYTES] */);
int gorder_comm_create(const uint8_t *id, int32_t n_ranks, int32_t rank, int32_t device, GorderComm **out);
int gorder_gpu_reduce_comm(GorderHandle *h, GorderComm *c, int32_t root);
/* Frequency::Once (leaflets.rs:435-441, 1523-1577): the table assigned from analysed frame 0 on rank `root` reaches the
 * other shards (collective; replaces the Arc<Mutex> shared by the reference's threads).  Single process: read the table from
 * the owning handle's results and pass it to gorder_gpu_set_leaflets. */
int gorder_comm_broadcast_leaflets(GorderHandle *h, GorderComm *c, int32_t root);
void gorder_comm_destroy(GorderComm *c);

/* Multi-GPU: the integer accumulators of this handle as ONE contiguous device block so that the
 * host can combine shards with a single NCCL sum-reduce. Layout: int64 words,
 *   [sum n_slots*3][count n_slots*3][map_sum n_slots*3*n_map_bins][map_count ...]
 * gorder_gpu_finish() reads the block back, so reduce first, then finish on the root. */
int gorder_gpu_accumulator_block(GorderHandle *h, void **d_ptr, int64_t *n_words);

/* Copy the accumulator block to / from caller-owned DEVICE memory (n_words int64 words): the host
 * reduces shards with  read -> ncclReduce(sum, int64) -> write on the root -> finish. */
int gorder_gpu_read_block(GorderHandle *h, void *d_dst);
int gorder_gpu_write_block(GorderHandle *h, const void *d_src);

/* Optional CUDA-event timing of the accumulation kernel (K1 / K2) on the handle's own stream.
 * profile_read returns the summed duration (ms) and the number of timed launches since the last
 * read and resets both. */
int gorder_gpu_profile(GorderHandle *h, int enable);
int gorder_gpu_profile_read(GorderHandle *h, double *hot_kernel_ms, int64_t *hot_kernel_launches);
/* The same for the membrane-normal stage of dynamic / manual normals (cell list + PCA kernels of a batch, normal.rs:160-199):
 * summed duration and number of batches since the last read. */
int gorder_gpu_profile_read_normals(GorderHandle *h, double *stage_ms, int64_t *batches);

/* Counters for benchmarking: kernels launched by this handle and frames analysed so far. */
int gorder_gpu_stats(GorderHandle *h, int64_t *kernel_launches, int64_t *frames);

/* Speculative Global leaflets (DESIGN.md §4, K1 SPEC): frames classified without a centre pre-pass and, of those,
 * frames that needed the exact two-pass centre afterwards (spec_repair_kernel).  `enabled` is 0 when the
 * configuration does not qualify or the engine switched the path off (too many repairs).  Results are identical
 * either way; the counters only explain throughput.  Blocks until the queued batches are done. */
int gorder_gpu_speculation_stats(GorderHandle *h, int32_t *enabled, int64_t *frames_speculated, int64_t *frames_repaired);

/* Asynchronous join: the main stream (gorder_gpu_stream) waits for the work this handle queued on its helper
 * streams (the tail of a batch -- repair + fold -- may run beside the next batch's kernel).  An event recorded on
 * the main stream after this call covers everything submitted so far.  Does not block the host. */
int gorder_gpu_fence(GorderHandle *h);

/* CUDA stream the accumulation kernels of this handle are queued on (as void*), for event timing by the caller
 * (call gorder_gpu_fence before recording the closing event). */
void *gorder_gpu_stream(GorderHandle *h);

/* ---- host-side trajectory feed ---------------------------------------------------------
 * Stands in for the reference's reader -- read_trajectory -> groan_rs traj_iter_map_reduce::<GroupXtcReader>
 * (src/analysis/common.rs:281-304) -> molly 0.5.0 -- when the harness needs a real end-to-end run from an .xtc file.
 * The Rust host keeps its own reader; these entry points need no GPU except gorder_gpu_run_xtc. */
typedef struct GorderXtc GorderXtc;
/* mmap the file and index its frames (headers only). */
int gorder_xtc_open(const char *path, GorderXtc **out);
int gorder_xtc_info(GorderXtc *x, int32_t *n_atoms, int64_t *n_frames, float *precision);
/* Decode frames first, first + stride, ... (count of them) with n_threads host threads into
 * xyz [count][n_atoms][3]; box9 [count][9], time [count], step [count] may be NULL. */
int gorder_xtc_read(GorderXtc *x, int64_t first, int64_t count, int64_t stride, int32_t n_threads, float *xyz, float *box9, float *time,
                    int32_t *step);
/* Write frames with orthogonal boxes (xdrfile's compression): xyz [n_frames][n_atoms][3], box3 [n_frames][3]. */
int gorder_xtc_write(const char *path, const float *xyz, const float *box3, int32_t n_atoms, int64_t n_frames, float precision, int32_t append,
                     int32_t first_step, float dt, int32_t n_threads);
void gorder_xtc_close(GorderXtc *x);
/* analyze_frame for frames first, first + stride, ... < last: n_threads host threads decode straight into pinned
 * batches in the plane layout (only the atoms the analysis needs), the decode of batch k+1 overlaps the copy and the
 * kernels of batch k.  atom_of_slot[s] = trajectory atom of engine atom s (NULL: identity).  The j-th analysed frame
 * gets frame_index frame_index0 + j * stride (topology/mod.rs:141-144; frame_index0 = 0, or the continuation value when
 * several files are concatenated, common.rs:306-339).  A non-orthogonal box fails with GORDER_ERR_NOT_ORTHOGONAL_BOX
 * (common.rs:186-198).  decode_seconds (optional): host time spent decoding, summed over threads. */
int gorder_gpu_run_xtc(GorderHandle *h, GorderXtc *x, const int32_t *atom_of_slot, int64_t first, int64_t last, int64_t stride,
                       int64_t frame_index0, int32_t n_threads, int32_t batch_frames, double *decode_seconds);

/* The same with the decode on the DEVICE: host threads only copy the compressed frames into pinned batches (4-5 B per
 * atom cross PCIe instead of 12); xtc_walk_kernel follows the control bits of every frame and bookmarks every 32nd group,
 * xtc_decode_kernel unpacks the groups in parallel into the engine's staging frames.  Coordinates are bit-identical to the
 * host decoder's.  Frames the device path does not cover (> 64 bits per triple, <= 9 atoms) take the host decoder; a stream
 * the walk cannot follow is skipped and reported as a deferred error (GORDER_ERR_INVALID_ARGUMENT, detail = the frame's
 * number in the file) by the next gorder_gpu_sync / gorder_gpu_finish.  batch_frames >= 128 hides the walk (a few ms per
 * batch whatever its size) behind the copies.  bytes_h2d (optional): bytes that crossed PCIe. */
int gorder_gpu_run_xtc_device(GorderHandle *h, GorderXtc *x, const int32_t *atom_of_slot, int64_t first, int64_t last, int64_t stride,
                              int64_t frame_index0, int32_t n_threads, int32_t batch_frames, int64_t *bytes_h2d);

/* ---- results conversion, the step after the path -------------------------------------------
 * Host only: no GPU, no handle.  Stands in for SystemTopology::convert (topology/mod.rs:122 ->
 * presentation/converter.rs:52-559) when the harness has no Rust converter: the reference's own converter keeps working on
 * the back-filled SystemTopology.  `slots` selects the accumulator slots that are summed before the division
 * (OrderSummer, converter.rs:513-559): one slot = one bond, the bonds of a heavy atom = the atom, all slots of a molecule
 * type = the molecule, all slots = the system.  Outputs are [total, upper, lower]; `sign` is -1 for AA/UA (-S_CH,
 * presentation/mod.rs:618-691) and +1 for CG. */
typedef struct GorderRaw {
    int32_t n_slots;
    int64_t n_frames;          /* rows of tw_* (0: no per-frame data) */
    const int64_t *sum;        /* [n_slots][3]   as filled by gorder_gpu_finish */
    const uint64_t *count;     /* [n_slots][3] */
    const int64_t *tw_sum;     /* [n_frames][n_slots][3] or NULL */
    const uint64_t *tw_count;  /* [n_frames][n_slots][3] or NULL */
} GorderRaw;
/* AnalysisOrder::calc_order (order.rs:97-107; NaN below min_samples) and, when `error` is not NULL and per-frame data
 * exist, TimeWiseData::estimate_error (timewise.rs:191-231): sample standard deviation of n_blocks block means (NaN when a
 * block is empty, when below min_samples or when there is no per-frame data). */
int gorder_results_order(const GorderRaw *raw, const int32_t *slots, int32_t n_sel, int32_t n_blocks, int32_t min_samples, float sign,
                         float *value /* [3] */, float *error /* [3] or NULL */);
/* TimeWiseData::prefix_average (timewise.rs:259-274): out[f][k] = order over frames 0..f (NaN while there are no samples). */
int gorder_results_convergence(const GorderRaw *raw, const int32_t *slots, int32_t n_sel, float sign, float *out /* [n_frames][3] */);
/* Order-map bins (ordermap.rs / converter.rs:159-308): out[i] = sign * (sum[i] / 1e6) / count[i], NaN below min_samples. */
int gorder_results_map(const int64_t *map_sum, const uint64_t *map_count, int64_t n, int32_t min_samples, float sign, float *out);

/* ---- structure, topology and molecule classification, the step before the path -----------------
 * Host only: no GPU, no handle.  Replaces, for a host without the Rust front end, read_structure_and_topology
 * (structure.rs:27-88: groan_rs System::from_file on a TPR -> minitpr 0.2.3; read_bonds :91-165) and
 * MoleculesClassifier::classify (topology/classify.rs:45-315, 318-580).  Atom groups are index lists: parsing the selection
 * language stays with the host.  gorder_topology_last_error() describes the last failure on the calling thread. */
typedef struct GorderSystem GorderSystem;
typedef struct GorderClassification GorderClassification;
const char *gorder_topology_last_error(void);
/* TPR of GROMACS 5.1 .. 2022 (tpx 103 .. 127, single or double precision): atoms (name, residue name / number, atomic number,
 * mass, charge), bonds (BONDS .. RESTRBONDS, CONSTR, CONSTRNC, the O-H pairs of SETTLE, intermolecular lists), box and
 * coordinates.  Anything else -> GORDER_ERR_TPR_FORMAT. */
int gorder_system_from_tpr(const char *path, GorderSystem **out);
/* read_structure_and_topology (structure.rs:27-88) by file extension: .tpr as above; .pdb with its CONECT records; .gro; a
 * bonds file (may be NULL) replaces the topology of any of them.  GORDER_ERR_NO_TOPOLOGY when no bonds come from anywhere,
 * GORDER_ERR_PDB_TOPOLOGY for CONECT records over repeated atom numbers, GORDER_ERR_STRUCTURE_FORMAT otherwise. */
int gorder_system_from_file(const char *structure, const char *bonds_file, GorderSystem **out);
/* A structure read by the host (GRO / PDB ...): names and residue numbers as arrays; xyz [n][3] and box9 may be NULL. */
int gorder_system_from_arrays(int32_t n_atoms, const char *const *atom_names, const char *const *res_names, const int32_t *res_ids,
                              const float *xyz, const float *box9, GorderSystem **out);
void gorder_system_free(GorderSystem *s);
int32_t gorder_system_n_atoms(const GorderSystem *s);
int64_t gorder_system_n_bonds(const GorderSystem *s);
int32_t gorder_system_tpx_version(const GorderSystem *s);   /* 0 unless read from a TPR */
/* names: [n][8] NUL-padded; any output may be NULL */
int gorder_system_atoms(const GorderSystem *s, char *atom_names8, char *res_names8, int32_t *res_ids, int32_t *atomic_numbers, float *masses,
                        float *charges);
int gorder_system_bonds(const GorderSystem *s, int32_t *pairs /* [n_bonds][2], i < j, sorted */);
int gorder_system_positions(const GorderSystem *s, float *xyz /* [n][3] */, int32_t *has);
int gorder_system_box(const GorderSystem *s, float *box9, int32_t *has);
/* Replace all bonds: by pairs of atom indices (from 0), or from a bonds file (structure.rs:91-165: `i j k ...`, serial
 * numbers from 1, '#' comments; GORDER_ERR_IO / _BONDS_PARSE / _BONDS_ATOM_NOT_FOUND / _BONDS_SELF as BondsError). */
int gorder_system_set_bonds(GorderSystem *s, const int32_t *pairs, int64_t n_pairs);
int gorder_system_read_bonds(GorderSystem *s, const char *bonds_file);
/* Molecule types for AA (group1 = heavy atoms, group2 = hydrogens) and CG (group1 = group2 = beads): classify.rs:53-76.
 * heads / methyls / normal_heads: the leaflet and dynamic-normal groups, NULL when unused (common.rs:345-375: exactly one
 * head per molecule; leaflets.rs:743-775: the same positive number of methyls in every molecule of a type).
 * As in the reference (classify.rs:297-315) a system without analysable molecules, or with a molecule type that has no order
 * bond, yields ZERO types and a warning, not an error. */
int gorder_classify_bonds(const GorderSystem *s, const int32_t *group1, int32_t n1, const int32_t *group2, int32_t n2,
                          const int32_t *heads, int32_t n_heads, const int32_t *methyls, int32_t n_methyls,
                          const int32_t *normal_heads, int32_t n_normal_heads, GorderClassification **out);
/* UA (classify.rs:77-90; carbon typing uaorder.rs:580-665; `ignore` = atoms that do not count as bonded). */
int gorder_classify_ua(const GorderSystem *s, const int32_t *saturated, int32_t n_sat, const int32_t *unsaturated, int32_t n_unsat,
                       const int32_t *ignore, int32_t n_ignore, const int32_t *heads, int32_t n_heads, const int32_t *methyls, int32_t n_methyls,
                       const int32_t *normal_heads, int32_t n_normal_heads, GorderClassification **out);
void gorder_classification_free(GorderClassification *c);
int32_t gorder_classification_n_types(const GorderClassification *c);
/* [n_types], ready for GorderSetup.moltypes (manual tables unset); valid until gorder_classification_free */
const GorderMolType *gorder_classification_moltypes(const GorderClassification *c);
const char *gorder_classification_type_name(const GorderClassification *c, int32_t type);   /* "POPC", "POPE-POPG", "POPC2" */
/* "POPC C22 (20) - POPC H2R (21)" for bond type i, "POPC C22 (20)" for united-atom carbon i */
const char *gorder_classification_item_name(const GorderClassification *c, int32_t type, int32_t item);
const char *gorder_classification_warning(const GorderClassification *c);   /* "" when molecule types were found */
int32_t gorder_classification_n_atoms_rel(const GorderClassification *c, int32_t type);
const int32_t *gorder_classification_atoms_rel(const GorderClassification *c, int32_t type);   /* relative indices of the molecule's atoms */

/* GROMACS index files (groan_rs Groups::from_ndx): `[ name ]` headers, atom numbers from 1.  A name with one of '"&|!@()<>= is
 * refused (its atoms are skipped), a repeated name replaces the earlier group; n_atoms < 0 skips the range check of the numbers. */
typedef struct GorderNdx GorderNdx;
int gorder_ndx_open(const char *path, int32_t n_atoms, GorderNdx **out);
void gorder_ndx_close(GorderNdx *n);
int32_t gorder_ndx_n_groups(const GorderNdx *n);
const char *gorder_ndx_group_name(const GorderNdx *n, int32_t group);
int64_t gorder_ndx_group_size(const GorderNdx *n, int32_t group);
const int32_t *gorder_ndx_group_atoms(const GorderNdx *n, int32_t group);   /* from 0, as listed */
int32_t gorder_ndx_find(const GorderNdx *n, const char *name);              /* -1: no such group */
/* LeafletClassification::FromNdx (leaflets.rs:1030-1215): one ndx file per assignment frame, groups `upper` / `lower` of head
 * atoms -> table[n_files][n_molecules] of GORDER_UPPER / GORDER_LOWER for the heads (absolute atom indices, molecule order) of
 * one molecule type: GorderMolType.manual_leaflets under GORDER_LEAFLET_MANUAL.  Errors as NdxLeafletClassificationError. */
int gorder_leaflets_from_ndx(const char *const *ndx_files, int32_t n_files, int32_t n_atoms, const char *upper, const char *lower,
                             const int32_t *heads, int32_t n_molecules, uint8_t *table);

/* The host stage of gorder_gpu_run_xtc_device without a GPU: walks the control bits of frames first .. first + count - 1
 * and reports per frame its groups (one "large" atom + its run of small ones) and the bookmarks the kernel would get.  A
 * cheap integrity check (nothing is decoded): n_groups[k] = -1 and GORDER_ERR_INVALID_ARGUMENT for an inconsistent stream;
 * n_groups[k] = -2 for a frame the device path hands to the host decoder (> 64 bits per small triple).  Either array may
 * be NULL. */
int gorder_xtc_scan(GorderXtc *x, int64_t first, int64_t count, int32_t *n_groups /* [count] */, int32_t *n_bookmarks /* [count] */);

/* Tuning hint: the smallest number of frames per batch for which the grid of the accumulation kernel (tiles x frames) is a
 * whole number of waves on this device; multiples of it avoid a partly filled last wave.  0 = no preference. */
int gorder_gpu_wave_frames(GorderHandle *h, int32_t *frames);

/* Human-readable detail of the last error of this handle (offending atom index etc.). */
int gorder_gpu_last_error(GorderHandle *h, char *buf, size_t len);

/* Offending index / count attached to the first deferred device error (e.g. atom slot). */
int64_t gorder_gpu_error_detail(GorderHandle *h);

void gorder_gpu_destroy(GorderHandle *h);

const char *gorder_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GORDER_B200_H */
