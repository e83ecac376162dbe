//! src/analysis/gpu.rs -- the shim a gorder maintainer adds INSIDE the `gorder` crate to run the per-frame
//! order-parameter engine on libgorder_b200.so (C ABI: include/gorder_b200.h).
//!
//! It replaces, for the AAOrder / CGOrder / UAOrder drivers (aaorder.rs:124, cgorder.rs:105, uaorder.rs:113),
//!   * `analyze_frame`                         (src/analysis/common.rs:201-235)  -> `gorder_gpu_submit`
//!   * `SystemTopology::new` + per-thread clone (topology/mod.rs:70-118, 274-277) -> `gorder_gpu_create`
//!   * `ParallelTrajData::reduce` + `Add` chain (topology/mod.rs:236-272)         -> `gorder_gpu_finish`
//!     (several GPUs: `gorder_gpu_reduce` / `gorder_gpu_reduce_comm` first)
//! and back-fills the crate-private accumulators (`AnalysisOrder`, `TimeWiseData`, `Map`, `AssignedLeaflets`,
//! `NormalsStorage`) so that `validate_run`, `error_info` and `convert` run unchanged.  It must live inside the crate
//! because those types are `pub(crate)`.
//!
//! The build image of this repository has no Rust toolchain: this file is not compiled here.  What IS checked
//! (tests/test_abi_cpu.py): every `#[repr(C)]` struct below has the fields of the C header, in order and with the
//! matching types, and every `extern "C"` function is exported by the library with the same argument count.
#![allow(non_camel_case_types, dead_code)]

use std::ffi::c_void;
use std::os::raw::c_char;

pub const GORDER_ABI_VERSION: i32 = 2;
pub const GORDER_COMM_ID_BYTES: usize = 128;

// ---- enums of the header (int32 in the structs) ----------------------------------------------------------------------
pub const GORDER_KIND_AA: i32 = 0;
pub const GORDER_KIND_CG: i32 = 1;
pub const GORDER_KIND_UA: i32 = 2;
pub const GORDER_NORMAL_STATIC: i32 = 0;
pub const GORDER_NORMAL_DYNAMIC: i32 = 1;
pub const GORDER_NORMAL_MANUAL: i32 = 2;
pub const GORDER_LEAFLET_NONE: i32 = 0;
pub const GORDER_LEAFLET_GLOBAL: i32 = 1;
pub const GORDER_LEAFLET_LOCAL: i32 = 2;
pub const GORDER_LEAFLET_INDIVIDUAL: i32 = 3;
pub const GORDER_LEAFLET_MANUAL: i32 = 4;
pub const GORDER_LEAFLET_SPHERICAL: i32 = 5;
pub const GORDER_FREQ_EVERY: i32 = 0;
pub const GORDER_FREQ_ONCE: i32 = 1;
pub const GORDER_GEOM_NONE: i32 = 0;
pub const GORDER_GEOM_CUBOID: i32 = 1;
pub const GORDER_GEOM_CYLINDER: i32 = 2;
pub const GORDER_GEOM_SPHERE: i32 = 3;
pub const GORDER_GEOMREF_POINT: i32 = 0;
pub const GORDER_GEOMREF_SELECTION: i32 = 1;
pub const GORDER_GEOMREF_BOX_CENTER: i32 = 2;
pub const GORDER_UA_CH3: i32 = 0;
pub const GORDER_UA_CH2: i32 = 1;
pub const GORDER_UA_CH1_UNSAT: i32 = 2;
pub const GORDER_UA_CH1_SAT: i32 = 3;

// error codes: 1:1 with AnalysisError (src/errors.rs:121-169) + device errors
pub const GORDER_OK: i32 = 0;
pub const GORDER_ERR_UNDEFINED_BOX: i32 = 1;
pub const GORDER_ERR_NOT_ORTHOGONAL_BOX: i32 = 2;
pub const GORDER_ERR_ZERO_BOX: i32 = 3;
pub const GORDER_ERR_UNDEFINED_POSITION: i32 = 4;
pub const GORDER_ERR_INVALID_GLOBAL_CENTER: i32 = 5;
pub const GORDER_ERR_INVALID_LOCAL_CENTER: i32 = 6;
pub const GORDER_ERR_MANUAL_LEAFLET_FRAME: i32 = 7;
pub const GORDER_ERR_DYNAMIC_NORMAL_POINTS: i32 = 8;
pub const GORDER_ERR_DYNAMIC_NORMAL_SVD: i32 = 9;
pub const GORDER_ERR_MANUAL_NORMAL_FRAME: i32 = 10;
pub const GORDER_ERR_LEAFLET_FRAME_UNAVAILABLE: i32 = 11;
pub const GORDER_ERR_ORDER_OVERFLOW: i32 = 12;
pub const GORDER_ERR_INVALID_ARGUMENT: i32 = 20;
pub const GORDER_ERR_NO_DEVICE: i32 = 30;
pub const GORDER_ERR_CUDA: i32 = 31;
pub const GORDER_ERR_OUT_OF_MEMORY: i32 = 32;
pub const GORDER_ERR_NCCL: i32 = 33;

// ---- structs: field for field as in include/gorder_b200.h -------------------------------------------------------------
#[repr(C)]
pub struct GorderMolType {
    pub n_molecules: i32,
    pub mol_base: *const i32,
    pub n_bond_types: i32,
    pub bond_rel: *const i32,
    pub n_ua_atoms: i32,
    pub ua_kind: *const i32,
    pub ua_rel: *const i32,
    pub head_rel: i32,
    pub n_methyls: i32,
    pub methyl_rel: *const i32,
    pub normal_head_rel: i32,
    pub n_manual_leaflet_frames: i32,
    pub manual_leaflets: *const u8,
    pub n_manual_normal_frames: i32,
    pub manual_normals: *const f32,
}

#[repr(C)]
pub struct GorderSetup {
    pub abi_version: i32,
    pub kind: i32,
    pub n_atoms: i32,
    pub handle_pbc: i32,
    pub step: i32,
    pub n_moltypes: i32,
    pub moltypes: *const GorderMolType,
    pub normal_mode: i32,
    pub normal_axis: i32,
    pub dynamic_radius: f32,
    pub n_normal_heads: i32,
    pub normal_heads: *const i32,
    pub collect_normals: i32,
    pub leaflet_mode: i32,
    pub leaflet_axis: i32,
    pub leaflet_freq_kind: i32,
    pub leaflet_freq: i32,
    pub leaflet_flip: i32,
    pub leaflet_radius: f32,
    pub n_membrane: i32,
    pub membrane: *const i32,
    pub collect_leaflets: i32,
    pub geom_kind: i32,
    pub geom_invert: i32,
    pub geom_ref_kind: i32,
    pub geom_ref_point: [f32; 3],
    pub n_geom_ref: i32,
    pub geom_ref: *const i32,
    pub geom_dims: [f32; 6],
    pub geom_axis: i32,
    pub structure_box: [f32; 3],
    pub map_enabled: i32,
    pub map_plane: i32,
    pub map_span_x: [f32; 2],
    pub map_span_y: [f32; 2],
    pub map_bin: [f32; 2],
    pub timewise: i32,
    pub device: i32,
    pub max_batch_frames: i32,
}

#[repr(C)]
pub struct GorderResults {
    pub n_slots: i64,
    pub n_frames: i64,
    pub n_map_bins: i64,
    pub map_nx: i64,
    pub map_ny: i64,
    pub n_leaflet_frames: i64,
    pub n_molecules_total: i64,
    pub sum: *mut i64,
    pub count: *mut u64,
    pub tw_sum: *mut i64,
    pub tw_count: *mut u64,
    pub tw_frame_index: *mut i64,
    pub map_sum: *mut i64,
    pub map_count: *mut u64,
    pub leaflets: *mut u8,
    pub leaflet_frame_index: *mut i64,
    pub normals: *mut f32,
}

#[repr(C)]
pub struct GorderHandle {
    _private: [u8; 0],
}
#[repr(C)]
pub struct GorderComm {
    _private: [u8; 0],
}

#[link(name = "gorder_b200")]
extern "C" {
    pub fn gorder_gpu_create(setup: *const GorderSetup, out: *mut *mut GorderHandle) -> i32;
    pub fn gorder_gpu_submit(h: *mut GorderHandle, xyz: *const f32, box_: *const f32, frame_index: *const i64, n_frames: i32) -> i32;
    pub fn gorder_gpu_native_layout(h: *mut GorderHandle, frame_floats: *mut i64, plane_offset: *mut i32, plane_cstride: *mut i32) -> i32;
    pub fn gorder_gpu_submit_native(h: *mut GorderHandle, planes: *const f32, box_: *const f32, frame_index: *const i64, n_frames: i32) -> i32;
    pub fn gorder_gpu_reserve_frames(h: *mut GorderHandle, n_frames: i64) -> i32;
    pub fn gorder_gpu_set_leaflets(h: *mut GorderHandle, table: *const u8, frame_index: i64) -> i32;
    pub fn gorder_gpu_sync(h: *mut GorderHandle) -> i32;
    pub fn gorder_gpu_result_sizes(h: *mut GorderHandle, r: *mut GorderResults) -> i32;
    pub fn gorder_gpu_finish(h: *mut GorderHandle, r: *mut GorderResults) -> i32;
    pub fn gorder_gpu_reduce(handles: *mut *mut GorderHandle, n: i32, root: i32) -> i32;
    pub fn gorder_comm_unique_id(id: *mut u8) -> i32;
    pub fn gorder_comm_create(id: *const u8, n_ranks: i32, rank: i32, device: i32, out: *mut *mut GorderComm) -> i32;
    pub fn gorder_gpu_reduce_comm(h: *mut GorderHandle, c: *mut GorderComm, root: i32) -> i32;
    pub fn gorder_comm_broadcast_leaflets(h: *mut GorderHandle, c: *mut GorderComm, root: i32) -> i32;
    pub fn gorder_comm_destroy(c: *mut GorderComm);
    pub fn gorder_gpu_last_error(h: *mut GorderHandle, buf: *mut c_char, len: usize) -> i32;
    pub fn gorder_gpu_error_detail(h: *mut GorderHandle) -> i64;
    pub fn gorder_gpu_destroy(h: *mut GorderHandle);
}

// ---- the shim proper (uses crate-private types: compiles only inside `gorder`) ----------------------------------------
#[cfg(feature = "gpu")]
mod shim {
    use super::*;
    use crate::analysis::common::check_box;
    use crate::analysis::topology::SystemTopology;
    use crate::errors::AnalysisError;
    use groan_rs::prelude::*;

    /// Frames per `gorder_gpu_submit`: 64 x 1M atoms x 12 B = 768 MB of pinned memory.
    const BATCH_FRAMES: usize = 64;

    /// Everything `GorderSetup` points to, kept alive until `gorder_gpu_create` has copied it.
    pub(crate) struct SetupOwner {
        pub c: GorderSetup,
        pub moltypes: Vec<GorderMolType>,
        pub mol_base: Vec<Vec<i32>>,
        pub bond_rel: Vec<Vec<i32>>,
        pub ua_kind: Vec<Vec<i32>>,
        pub ua_rel: Vec<Vec<i32>>,
        pub methyl_rel: Vec<Vec<i32>>,
        pub membrane: Vec<i32>,
        pub normal_heads: Vec<i32>,
        pub geom_ref: Vec<i32>,
        /// absolute atom index -> slot of the frame handed to the engine (the Master group in index order, common.rs:92-103)
        pub slot_of: Vec<i32>,
    }

    /// Decoded frames of one batch: [frame][slot][xyz] f32 (NaN = no position), boxes, analysed-frame indices.
    pub(crate) struct Collector {
        xyz: Vec<f32>,
        box_: Vec<f32>,
        idx: Vec<i64>,
        n_slots: usize,
        pub frame: usize,
        pub step: usize,
        pub n_threads: usize,
    }

    impl Collector {
        /// Copy the Master-group positions of `frame` into the batch (replaces the body of analyze_frame, common.rs:201-235).
        pub(crate) fn push(&mut self, frame: &System, master: &[usize]) -> Result<(), AnalysisError> {
            let simbox = check_box(frame)?; // UndefinedBox / NotOrthogonalBox / ZeroBox stay on the host (common.rs:186-198)
            for &a in master {
                match frame.get_atom(a).unwrap().get_position() {
                    Some(p) => self.xyz.extend_from_slice(&[p.x, p.y, p.z]),
                    None => self.xyz.extend_from_slice(&[f32::NAN; 3]), // -> GORDER_ERR_UNDEFINED_POSITION(slot)
                }
            }
            self.box_.extend_from_slice(&[simbox.x, simbox.y, simbox.z]);
            self.idx.push(self.frame as i64);
            self.frame += self.step * self.n_threads; // topology/mod.rs:141-144
            Ok(())
        }
        pub(crate) fn len(&self) -> usize {
            self.idx.len()
        }
        pub(crate) fn flush(&mut self, h: *mut GorderHandle) -> Result<(), AnalysisError> {
            if self.idx.is_empty() {
                return Ok(());
            }
            let rc = unsafe { gorder_gpu_submit(h, self.xyz.as_ptr(), self.box_.as_ptr(), self.idx.as_ptr(), self.idx.len() as i32) };
            self.xyz.clear();
            self.box_.clear();
            self.idx.clear();
            check(rc, h)
        }
    }

    /// Return code -> AnalysisError of the same name; the detail is the atom index / frame / point count the variant carries.
    pub(crate) fn check(rc: i32, h: *mut GorderHandle) -> Result<(), AnalysisError> {
        if rc == GORDER_OK {
            return Ok(());
        }
        let detail = unsafe { gorder_gpu_error_detail(h) };
        Err(match rc {
            GORDER_ERR_UNDEFINED_BOX => AnalysisError::UndefinedBox,
            GORDER_ERR_NOT_ORTHOGONAL_BOX => AnalysisError::NotOrthogonalBox,
            GORDER_ERR_ZERO_BOX => AnalysisError::ZeroBox,
            GORDER_ERR_UNDEFINED_POSITION => AnalysisError::UndefinedPosition(detail as usize),
            GORDER_ERR_INVALID_GLOBAL_CENTER => AnalysisError::InvalidGlobalMembraneCenter,
            GORDER_ERR_INVALID_LOCAL_CENTER => AnalysisError::InvalidLocalMembraneCenter(detail as usize),
            _ => AnalysisError::Gpu(rc, last_error(h)), // one new variant: device / NCCL / argument errors with the library's message
        })
    }

    fn last_error(h: *mut GorderHandle) -> String {
        let mut buf = vec![0 as c_char; 512];
        unsafe { gorder_gpu_last_error(h, buf.as_mut_ptr(), buf.len()) };
        unsafe { std::ffi::CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned()
    }

    /// Arrays of `GorderResults`, allocated from `gorder_gpu_result_sizes`.
    pub(crate) struct ResultsOwner {
        pub c: GorderResults,
        pub sum: Vec<i64>,
        pub count: Vec<u64>,
        pub tw_sum: Vec<i64>,
        pub tw_count: Vec<u64>,
        pub tw_frame_index: Vec<i64>,
        pub map_sum: Vec<i64>,
        pub map_count: Vec<u64>,
        pub leaflets: Vec<u8>,
        pub leaflet_frame_index: Vec<i64>,
        pub normals: Vec<f32>,
    }

    /// `read_trajectory` (common.rs:239-342) with the engine behind it.  `build_setup` and `backfill` follow the two tables of
    /// INTEGRATION.md §2 / §3 (they only read / write fields of `SystemTopology`).
    pub(crate) fn read_trajectory_gpu(
        system: &mut System,
        mut topology: SystemTopology,
        trajectory: &str,
        n_threads: usize,
        begin: f32,
        end: f32,
        step: usize,
    ) -> Result<SystemTopology, Box<dyn std::error::Error + Send + Sync>> {
        let setup = build_setup(system, &topology)?;
        let mut h: *mut GorderHandle = std::ptr::null_mut();
        let rc = unsafe { gorder_gpu_create(&setup.c, &mut h) };
        if rc != GORDER_OK {
            return Err(Box::new(AnalysisError::Gpu(rc, "gorder_gpu_create failed (no CPU fallback)".into())));
        }
        let master: Vec<usize> = system.group_iter("Master")?.map(|a| a.get_index()).collect();
        let mut collector = Collector {
            xyz: Vec::with_capacity(BATCH_FRAMES * master.len() * 3),
            box_: Vec::new(),
            idx: Vec::new(),
            n_slots: master.len(),
            frame: 0,
            step,
            n_threads: 1, // the decode threads hand their frames over in trajectory order; the engine owns the parallelism
        };
        for frame in system.group_xtc_iter(trajectory, "Master")?.with_range(begin, end)?.with_step(step)? {
            let frame = frame?;
            collector.push(frame, &master)?;
            if collector.len() == BATCH_FRAMES {
                collector.flush(h)?;
            }
        }
        collector.flush(h)?;
        let mut r = ResultsOwner::sized(h, &setup.c)?;
        check(unsafe { gorder_gpu_finish(h, &mut r.c) }, h)?; // replaces ParallelTrajData::reduce
        backfill(&mut topology, &r, &setup);
        unsafe { gorder_gpu_destroy(h) };
        Ok(topology) // validate_run / error_info / convert unchanged
    }

    impl ResultsOwner {
        /// `gorder_gpu_result_sizes`, then one Vec per array the run produces (an array left NULL is skipped by the library).
        pub(crate) fn sized(h: *mut GorderHandle, setup: &GorderSetup) -> Result<ResultsOwner, AnalysisError> {
            let mut c: GorderResults = unsafe { std::mem::zeroed() };
            check(unsafe { gorder_gpu_result_sizes(h, &mut c) }, h)?;
            let (ns, nf, nb) = (c.n_slots as usize, c.n_frames as usize, c.n_map_bins as usize);
            let (nl, nm) = (c.n_leaflet_frames as usize, c.n_molecules_total as usize);
            let tw = setup.timewise != 0;
            let mut r = ResultsOwner {
                sum: vec![0; ns * 3],
                count: vec![0; ns * 3],
                tw_sum: vec![0; if tw { nf * ns * 3 } else { 0 }],
                tw_count: vec![0; if tw { nf * ns * 3 } else { 0 }],
                tw_frame_index: vec![0; nf],
                map_sum: vec![0; ns * 3 * nb],
                map_count: vec![0; ns * 3 * nb],
                leaflets: vec![0; nl * nm],
                leaflet_frame_index: vec![0; nl],
                normals: vec![f32::NAN; if setup.collect_normals != 0 { nf * nm * 3 } else { 0 }],
                c,
            };
            fn p<T>(v: &mut Vec<T>) -> *mut T {
                if v.is_empty() { std::ptr::null_mut() } else { v.as_mut_ptr() }
            }
            r.c.sum = p(&mut r.sum);
            r.c.count = p(&mut r.count);
            r.c.tw_sum = p(&mut r.tw_sum);
            r.c.tw_count = p(&mut r.tw_count);
            r.c.tw_frame_index = p(&mut r.tw_frame_index);
            r.c.map_sum = p(&mut r.map_sum);
            r.c.map_count = p(&mut r.map_count);
            r.c.leaflets = p(&mut r.leaflets);
            r.c.leaflet_frame_index = p(&mut r.leaflet_frame_index);
            r.c.normals = p(&mut r.normals);
            Ok(r)
        }
    }

    /// `GorderSetup` from the classified topology.  Molecules of one type are congruent
    /// (topology/molecule.rs:224-244): the relative indices of the FIRST instance describe all of them; the shim asserts it.
    fn build_setup(system: &System, topology: &SystemTopology) -> Result<SetupOwner, AnalysisError> {
        let master: Vec<usize> = system.group_iter("Master").unwrap().map(|a| a.get_index()).collect();
        let mut slot_of = vec![-1i32; system.get_n_atoms()];
        for (s, &a) in master.iter().enumerate() {
            slot_of[a] = s as i32;
        }
        let mut o = SetupOwner::empty(slot_of);
        for mol in topology.molecule_types().iter_bonds() {
            // AA / CG: bond types in the sorted order of OrderBonds::new (bond.rs:77-81)
            let bonds = mol.order_structure().bond_types();
            let first = bonds.first().expect("molecule type without bonds");
            let n_mol = first.bonds().len();
            // min_index of every instance = smallest absolute index over the instance's atoms; relative index = abs - min_index
            let rel = |abs: usize, m: usize| abs - mol.min_index_of_instance(m);
            let mut base = Vec::with_capacity(n_mol);
            for m in 0..n_mol {
                base.push(o.slot_of[mol.min_index_of_instance(m)]);
            }
            let mut pairs = Vec::with_capacity(bonds.len() * 2);
            for b in bonds {
                let (a1, a2) = b.bonds()[0];
                pairs.push(rel(a1, 0) as i32);
                pairs.push(rel(a2, 0) as i32);
                for (m, &(x1, x2)) in b.bonds().iter().enumerate() {
                    assert_eq!((rel(x1, m), rel(x2, m)), (rel(a1, 0), rel(a2, 0)), "molecules of a type must be congruent");
                }
            }
            o.push_moltype(base, pairs, mol);
        }
        o.fill_groups(system, topology); // Membrane, NormalHeads, GeomReference -> slots; leaflet / normal / geometry / map parameters
        o.c.abi_version = GORDER_ABI_VERSION;
        o.c.n_atoms = master.len() as i32;
        Ok(o)
    }

    /// Integer accumulators -> the crate's private structures: afterwards the reference's own
    /// `validate_run`, `error_info` and `convert` see exactly what their CPU frame loop would have left.
    fn backfill(topology: &mut SystemTopology, r: &ResultsOwner, setup: &SetupOwner) {
        let nf = r.c.n_frames as usize;
        let ns = r.c.n_slots as usize;
        let nb = r.c.n_map_bins as usize;
        let mut slot = 0usize;
        for mol in topology.molecule_types_mut().iter_bonds_mut() {
            for bond in mol.order_structure_mut().bond_types_mut().iter_mut() {
                for (k, acc) in [Some(bond.total_mut()), bond.upper_mut().as_mut(), bond.lower_mut().as_mut()].into_iter().enumerate() {
                    let Some(acc) = acc else { continue };
                    // AnalysisOrder { order: OrderValue(i64), n_samples, timewise } (order.rs:70-79)
                    acc.set_raw(r.sum[slot * 3 + k], r.count[slot * 3 + k] as usize);
                    if setup.c.timewise != 0 {
                        // TimeWiseData { order, n_samples, n_threads: 1 } (timewise.rs:131-140): rows are already in frame order
                        acc.set_timewise_raw(
                            (0..nf).map(|f| r.tw_sum[(f * ns + slot) * 3 + k]),
                            (0..nf).map(|f| r.tw_count[(f * ns + slot) * 3 + k] as usize),
                        );
                    }
                }
                if nb > 0 {
                    // Map.values / Map.samples (ordermap.rs:21-31): x-major, as GridMap::extract_raw walks them
                    for (k, map) in [bond.total_map_mut().as_mut(), bond.upper_map_mut().as_mut(), bond.lower_map_mut().as_mut()].into_iter().enumerate() {
                        let Some(map) = map else { continue };
                        let at = (slot * 3 + k) * nb;
                        map.set_raw(&r.map_sum[at..at + nb], &r.map_count[at..at + nb]);
                    }
                }
                slot += 1;
            }
        }
        topology.set_total_frames(nf); // topology/mod.rs:49
        // AssignedLeaflets.shared for the export (leaflets.rs:1371-1381, 1523): 1 = upper, 0 = lower (lib.rs:416-422)
        topology.set_collected_leaflets(&r.leaflet_frame_index, &r.leaflets, r.c.n_molecules_total as usize);
        // NormalsStorage.normals (normal.rs:462-467): [frame][molecule] Vector3D, NaN where no normal was requested
        topology.set_collected_normals(&r.normals, nf, r.c.n_molecules_total as usize);
    }
}
