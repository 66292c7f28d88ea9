//! `StarkV1Cuda`: the reference's `ProvingBackend` and `ProvingBackendStream` over libsezkp_cuda.so
//! (include/sezkp_cuda.h).  Source only — not compiled in this repository's image (no Rust toolchain); kept in sync
//! with INTEGRATION.md §2.  The same C entry points are exercised by tools/c_driver.c (plain C) and the Python mirror.
//!
//! One library context lives for the whole process (`OnceLock`): device pools, twiddle / coset tables and the pinned
//! staging ring are built once, not per `prove`.  With more than one GPU visible (or `SEZKP_CUDA_DEVICES=0,1,..`) the
//! context is a multi-GPU group (`sezkp_cuda_create_multi`): every `prove` / stream then shards one proof over all of
//! them inside the library — the Rust side stays the reference's single-process, synchronous call.
use anyhow::{anyhow, bail, ensure, Result};
use sezkp_core::{BackendKind, BlockSummary, ProofArtifact, ProvingBackend, ProvingBackendStream};
use std::ffi::{c_char, c_void, CStr};
use std::sync::{Mutex, OnceLock};

#[repr(C)]
pub struct TraceDesc {            // include/sezkp_trace.h
    tau: u32, flags: u32, n_blocks: u64, n_rows: u64,
    block_len: *const u64, win_left: *const i64, win_right: *const i64,
    head_in_off: *const u32, head_out_off: *const u32,
    input_mv: *const i8, mv: *const i8, write_flag: *const u8, write_sym: *const u16,
}
#[allow(non_camel_case_types)] type sezkp_ctx = c_void;
#[allow(non_camel_case_types)] type sezkp_stream = c_void;
extern "C" {
    fn sezkp_cuda_abi_version() -> u32;
    fn sezkp_cuda_device_count() -> i32;
    fn sezkp_cuda_create(device_id: i32, out: *mut *mut sezkp_ctx) -> i32;
    fn sezkp_cuda_create_multi(device_ids: *const i32, n_dev: i32, out: *mut *mut sezkp_ctx) -> i32;
    fn sezkp_cuda_last_error(ctx: *const sezkp_ctx) -> *const c_char;
    fn sezkp_stark_v1_prove(ctx: *mut sezkp_ctx, trace: *const TraceDesc, manifest_root: *const u8,
                            proof_buf: *mut u8, cap: usize, len: *mut usize) -> i32;
    fn sezkp_stark_v1_proof_bound(n_rows: u64, tau: u32) -> usize;
    fn sezkp_stark_v1_begin(ctx: *mut sezkp_ctx, tau: u32, manifest_root: *const u8, expected_rows: u64,
                            out: *mut *mut sezkp_stream) -> i32;
    fn sezkp_stark_v1_ingest(ctx: *mut sezkp_ctx, st: *mut sezkp_stream, blocks: *const TraceDesc) -> i32;
    fn sezkp_stark_v1_finish(ctx: *mut sezkp_ctx, st: *mut sezkp_stream, proof_buf: *mut u8, cap: usize, len: *mut usize) -> i32;
    fn sezkp_stark_v1_abort(ctx: *mut sezkp_ctx, st: *mut sezkp_stream);
}

/// The process-wide library context.  A ctx is not thread-safe (one call at a time), hence the mutex; it is never
/// destroyed (the reference's backends are stateless associated functions with no shutdown hook).
struct Ctx(*mut sezkp_ctx);
unsafe impl Send for Ctx {}
static CTX: OnceLock<std::result::Result<Mutex<Ctx>, String>> = OnceLock::new();

fn last_error(ctx: *const sezkp_ctx) -> String {
    unsafe { CStr::from_ptr(sezkp_cuda_last_error(ctx)).to_string_lossy().into_owned() }
}
fn with_ctx<T>(f: impl FnOnce(*mut sezkp_ctx) -> Result<T>) -> Result<T> {
    let slot = CTX.get_or_init(|| unsafe {
        if sezkp_cuda_abi_version() != 1 { return Err("libsezkp_cuda ABI mismatch".into()); }
        let devices: Vec<i32> = match std::env::var("SEZKP_CUDA_DEVICES") {
            Ok(s) => s.split(',').filter_map(|x| x.trim().parse().ok()).collect(),
            Err(_) => (0..sezkp_cuda_device_count()).collect(),
        };
        let mut ctx = std::ptr::null_mut();
        let rc = if devices.len() > 1 { sezkp_cuda_create_multi(devices.as_ptr(), devices.len() as i32, &mut ctx) }
                 else { sezkp_cuda_create(devices.first().copied().unwrap_or(-1), &mut ctx) };
        if rc != 0 { return Err(format!("sezkp_cuda error {rc}: {}", last_error(std::ptr::null()))); }
        Ok(Mutex::new(Ctx(ctx)))
    });
    let guard = slot.as_ref().map_err(|e| anyhow!("{e}"))?.lock().map_err(|_| anyhow!("sezkp_cuda context poisoned"))?;
    f(guard.0)
}
fn check(ctx: *mut sezkp_ctx, rc: i32) -> Result<()> {
    if rc != 0 { bail!("sezkp_cuda error {rc}: {}", last_error(ctx)); }
    Ok(())
}

/// Flat copies of the fields prove_v1 reads (v1/columns.rs:252-365).
#[derive(Default)]
struct Flat { block_len: Vec<u64>, wl: Vec<i64>, wr: Vec<i64>, io: Vec<u32>, oo: Vec<u32>,
              imv: Vec<i8>, mv: Vec<i8>, wf: Vec<u8>, ws: Vec<u16>, tau: u32 }
impl Flat {
    fn push(&mut self, b: &BlockSummary) -> Result<()> {
        let tau = b.windows.len();
        if self.block_len.is_empty() { self.tau = tau as u32; }
        ensure!(tau as u32 == self.tau, "tau mismatch");
        let len = (b.step_hi - b.step_lo + 1) as usize;
        ensure!(b.movement_log.steps.len() == len, "movement log length != step range");
        self.block_len.push(len as u64);
        for r in 0..tau {
            self.wl.push(b.windows[r].left); self.wr.push(b.windows[r].right);
            self.io.push(b.head_in_offsets[r]); self.oo.push(b.head_out_offsets[r]);
        }
        for s in &b.movement_log.steps {
            self.imv.push(s.input_mv);
            for op in &s.tapes { self.mv.push(op.mv); self.wf.push(op.write.is_some() as u8); self.ws.push(op.write.unwrap_or(0)); }
        }
        Ok(())
    }
    fn desc(&self) -> TraceDesc {
        TraceDesc { tau: self.tau, flags: 0, n_blocks: self.block_len.len() as u64, n_rows: self.imv.len() as u64,
            block_len: self.block_len.as_ptr(), win_left: self.wl.as_ptr(), win_right: self.wr.as_ptr(),
            head_in_off: self.io.as_ptr(), head_out_off: self.oo.as_ptr(), input_mv: self.imv.as_ptr(), mv: self.mv.as_ptr(),
            write_flag: self.wf.as_ptr(), write_sym: self.ws.as_ptr() }
    }
}
fn artifact(manifest_root: [u8; 32], bytes: Vec<u8>, n_rows: usize, tau: u32, streaming: bool) -> ProofArtifact {
    let meta = if streaming { serde_json::json!({"proto": "stark-v1", "mode": "streaming", "domain_n": n_rows * 8, "tau": tau}) }
               else { serde_json::json!({"proto": "stark-v1", "domain_n": n_rows * 8, "tau": tau}) };
    ProofArtifact { backend: BackendKind::Stark, manifest_root, proof_bytes: bytes, meta }   // sezkp-stark/src/lib.rs:129-142, 170-190
}

pub struct StarkV1Cuda;
impl ProvingBackend for StarkV1Cuda {
    fn prove(blocks: &[BlockSummary], manifest_root: [u8; 32]) -> Result<ProofArtifact> {
        let mut f = Flat::default();
        for b in blocks { f.push(b)?; }
        let d = f.desc();
        with_ctx(|ctx| unsafe {
            let mut bytes = vec![0u8; sezkp_stark_v1_proof_bound(d.n_rows, d.tau)];
            let mut len = 0usize;
            check(ctx, sezkp_stark_v1_prove(ctx, &d, manifest_root.as_ptr(), bytes.as_mut_ptr(), bytes.len(), &mut len))?;
            bytes.truncate(len);
            Ok(artifact(manifest_root, bytes, f.imv.len(), f.tau, false))
        })
    }
    fn verify(a: &ProofArtifact, blocks: &[BlockSummary], root: [u8; 32]) -> Result<()> {
        sezkp_stark::StarkV1::verify(a, blocks, root)              // unchanged CPU verifier
    }
}

/// Push API (sezkp-core/src/prover.rs:21-33), driven by `StreamingProver::prove_stream_iter` (prover.rs:104-150): every
/// block is flattened and handed to the library, which packs it into pinned staging buffers and copies 2^20-row slabs to
/// the GPU on a side stream while the caller parses the next block.  The library stream is opened at the first block
/// (tau is not known before).
pub struct CudaStreamState { root: [u8; 32], st: *mut sezkp_stream, rows: usize, tau: u32 }
impl Drop for CudaStreamState {
    fn drop(&mut self) {                                           // abandoned stream: release device trace + staging ring
        if !self.st.is_null() { let st = self.st; let _ = with_ctx(|ctx| unsafe { sezkp_stark_v1_abort(ctx, st); Ok(()) }); }
    }
}
impl ProvingBackendStream for StarkV1Cuda {
    type StreamState = CudaStreamState;
    fn begin_stream(manifest_root: [u8; 32]) -> Result<CudaStreamState> {
        Ok(CudaStreamState { root: manifest_root, st: std::ptr::null_mut(), rows: 0, tau: 0 })
    }
    fn ingest_block(state: &mut CudaStreamState, block: BlockSummary) -> Result<()> {
        let mut f = Flat::default();
        f.push(&block)?;
        let d = f.desc();
        with_ctx(|ctx| unsafe {
            if state.st.is_null() {
                state.tau = f.tau;
                check(ctx, sezkp_stark_v1_begin(ctx, f.tau, state.root.as_ptr(), 0, &mut state.st))?;
            }
            ensure!(f.tau == state.tau, "tau mismatch");
            check(ctx, sezkp_stark_v1_ingest(ctx, state.st, &d))?;    // a rejected block leaves the stream as it was
            state.rows += f.imv.len();
            Ok(())
        })
    }
    fn finish_stream(mut state: CudaStreamState) -> Result<ProofArtifact> {
        ensure!(!state.st.is_null(), "no blocks were ingested");
        with_ctx(|ctx| unsafe {
            let mut bytes = vec![0u8; sezkp_stark_v1_proof_bound(state.rows as u64, state.tau)];
            let mut len = 0usize;
            check(ctx, sezkp_stark_v1_finish(ctx, state.st, bytes.as_mut_ptr(), bytes.len(), &mut len))?;  // consumes the handle on success
            state.st = std::ptr::null_mut();
            bytes.truncate(len);
            Ok(artifact(state.root, bytes, state.rows, state.tau, true))
        })
    }
}
