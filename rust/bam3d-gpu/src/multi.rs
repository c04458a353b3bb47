//! Multi-GPU (one process per GPU): pair lists shard across ranks with no data-path collective; the one exchange step is the
//! gather of the ranks' compacted contacts over NCCL (`bam3d_comm_*`, `bam3d_gather_contacts*`). The reference has no
//! distributed layer: this is the `gpu` feature's addition. The 256-byte rendezvous id is made by rank 0 and handed to the other
//! ranks by whatever channel the application has (MPI, a socket, a file).
use crate::{ffi, Context, Error, Result};

pub struct Comm { pub(crate) h: *mut ffi::Comm, pub rank: i32, pub n_ranks: i32 }
impl Comm {
    pub fn unique_id() -> Result<[u8; ffi::UNIQUE_ID_BYTES]> {
        let mut id = [0u8; ffi::UNIQUE_ID_BYTES];
        let rc = unsafe { ffi::bam3d_comm_unique_id(id.as_mut_ptr()) };
        if rc == ffi::OK { Ok(id) } else { Err(Error { code: rc, message: "bam3d_comm_unique_id: libnccl.so.2 could not be loaded".into() }) }
    }
    /// Collective: every rank calls it with the same id.
    pub fn new(device: i32, id: &[u8; ffi::UNIQUE_ID_BYTES], rank: i32, n_ranks: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::bam3d_comm_create(device, id.as_ptr(), rank, n_ranks, &mut h) };
        if rc == ffi::OK { Ok(Comm { h, rank, n_ranks }) } else { Err(Error { code: rc, message: "bam3d_comm_create failed".into() }) }
    }
    /// This rank's dense results (as `GJK::intersection_batch` returned them for its slice, whose global indices start at
    /// `pair_index_base`) in, the hits of ALL ranks out, in rank order.
    pub fn gather_contacts(&self, ctx: &Context, contacts: &[ffi::RawContact], status: &[u8], pair_index_base: u64, capacity: usize)
                           -> Result<(Vec<ffi::RawContact40>, Vec<u64>)> {
        let (mut out, mut counts, mut total) = (vec![ffi::RawContact40::default(); capacity.max(1)], vec![0u64; self.n_ranks as usize], 0u64);
        ctx.check(unsafe { ffi::bam3d_gather_contacts(ctx.0, self.h, contacts.as_ptr(), status.as_ptr(), status.len() as u64, pair_index_base, out.as_mut_ptr(),
                                                      capacity as u64, counts.as_mut_ptr(), &mut total) })?;
        out.truncate(total as usize);
        Ok((out, counts))
    }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { ffi::bam3d_comm_destroy(self.h) } } }
