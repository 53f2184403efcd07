//! Raw bindings to `libzksc` -- see `include/zksc.h` for the contract of every entry point.
//! Field elements are BLS12-381 Fr in ark-ff's in-memory form: 4 x u64 little-endian limbs, Montgomery (R = 2^256).
mod ffi;
pub use ffi::*;
