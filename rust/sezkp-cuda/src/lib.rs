//! `StarkV1Cuda`: the reference's `ProvingBackend` over libsezkp_cuda.so (include/sezkp_cuda.h).
//! Source only — not compiled in this repository's image (no Rust toolchain); kept in sync with INTEGRATION.md §2.
use anyhow::{bail, ensure, Result};
use sezkp_core::{BackendKind, BlockSummary, ProofArtifact, ProvingBackend};
use std::ffi::{c_char, c_void, CStr};

#[repr(C)]
pub struct TraceDesc {            // include/sezkp_trace.h
    tau: u32, flags: u32, n_blocks: u64, n_rows: u64,
    block_len: *const u64, win_left: *const i64, win_right: *const i64,
    head_in_off: *const u32, head_out_off: *const u32,
    input_mv: *const i8, mv: *const i8, write_flag: *const u8, write_sym: *const u16,
}
#[allow(non_camel_case_types)] type sezkp_ctx = c_void;
extern "C" {
    fn sezkp_cuda_abi_version() -> u32;
    fn sezkp_cuda_create(device_id: i32, out: *mut *mut sezkp_ctx) -> i32;
    fn sezkp_cuda_destroy(ctx: *mut sezkp_ctx);
    fn sezkp_cuda_last_error(ctx: *const sezkp_ctx) -> *const c_char;
    fn sezkp_stark_v1_prove(ctx: *mut sezkp_ctx, trace: *const TraceDesc, manifest_root: *const u8,
                            proof_buf: *mut u8, cap: usize, len: *mut usize) -> i32;
}

/// Flat copies of the fields prove_v1 reads (v1/columns.rs:252-365).
struct Flat { block_len: Vec<u64>, wl: Vec<i64>, wr: Vec<i64>, io: Vec<u32>, oo: Vec<u32>,
              imv: Vec<i8>, mv: Vec<i8>, wf: Vec<u8>, ws: Vec<u16>, tau: u32 }
fn flatten(blocks: &[BlockSummary]) -> Result<Flat> {
    let tau = blocks.first().map(|b| b.windows.len()).unwrap_or(0);
    let mut f = Flat { block_len: vec![], wl: vec![], wr: vec![], io: vec![], oo: vec![], imv: vec![], mv: vec![],
                       wf: vec![], ws: vec![], tau: tau as u32 };
    for b in blocks {
        ensure!(b.windows.len() == tau, "tau mismatch");
        let len = (b.step_hi - b.step_lo + 1) as usize;
        ensure!(b.movement_log.steps.len() == len, "movement log length != step range");
        f.block_len.push(len as u64);
        for r in 0..tau {
            f.wl.push(b.windows[r].left); f.wr.push(b.windows[r].right);
            f.io.push(b.head_in_offsets[r]); f.oo.push(b.head_out_offsets[r]);
        }
        for s in &b.movement_log.steps {
            f.imv.push(s.input_mv);
            for op in &s.tapes { f.mv.push(op.mv); f.wf.push(op.write.is_some() as u8); f.ws.push(op.write.unwrap_or(0)); }
        }
    }
    Ok(f)
}

pub struct StarkV1Cuda;
impl ProvingBackend for StarkV1Cuda {
    fn prove(blocks: &[BlockSummary], manifest_root: [u8; 32]) -> Result<ProofArtifact> {
        let f = flatten(blocks)?;
        let d = TraceDesc { tau: f.tau, flags: 0, n_blocks: f.block_len.len() as u64, n_rows: f.imv.len() as u64,
            block_len: f.block_len.as_ptr(), win_left: f.wl.as_ptr(), win_right: f.wr.as_ptr(),
            head_in_off: f.io.as_ptr(), head_out_off: f.oo.as_ptr(), input_mv: f.imv.as_ptr(), mv: f.mv.as_ptr(),
            write_flag: f.wf.as_ptr(), write_sym: f.ws.as_ptr() };
        unsafe {
            ensure!(sezkp_cuda_abi_version() == 1, "libsezkp_cuda ABI mismatch");
            let mut ctx = std::ptr::null_mut();
            if sezkp_cuda_create(-1, &mut ctx) != 0 {
                bail!("{}", CStr::from_ptr(sezkp_cuda_last_error(std::ptr::null())).to_string_lossy());
            }
            let mut len = 0usize;                                   // two-call pattern: size, then fill
            let mut rc = sezkp_stark_v1_prove(ctx, &d, manifest_root.as_ptr(), std::ptr::null_mut(), 0, &mut len);
            let mut bytes = vec![0u8; len];
            if rc == 0 { rc = sezkp_stark_v1_prove(ctx, &d, manifest_root.as_ptr(), bytes.as_mut_ptr(), len, &mut len); }
            let err = if rc != 0 { Some(CStr::from_ptr(sezkp_cuda_last_error(ctx)).to_string_lossy().into_owned()) } else { None };
            sezkp_cuda_destroy(ctx);
            if let Some(e) = err { bail!("sezkp_cuda error {rc}: {e}"); }
            bytes.truncate(len);
            Ok(ProofArtifact { backend: BackendKind::Stark, manifest_root, proof_bytes: bytes,
                meta: serde_json::json!({"proto": "stark-v1", "domain_n": (f.imv.len() * 8), "tau": f.tau}) })
        }
    }
    fn verify(a: &ProofArtifact, blocks: &[BlockSummary], root: [u8; 32]) -> Result<()> {
        sezkp_stark::StarkV1::verify(a, blocks, root)              // unchanged CPU verifier
    }
}
