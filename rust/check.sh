#!/bin/sh
# Build the two crates and run the reference's protocol tests against the GPU shim.  Needs: cargo, nvcc (CUDA >= 12.9), a B200, and
# the reference checked out next to this repository (../zk-cryptography) with the one-line accessor of INTEGRATION.md section 1.
# NOT run in this repository's build image (no Rust toolchain there): rust/README.md.
set -eu
cd "$(dirname "$0")"
( cd zksc-sys && cargo build --release )
( cd zksc-sumcheck && cargo test --release -- --test-threads 1 )   # one context per thread: keep the GPU tests on one thread
