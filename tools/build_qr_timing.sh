#!/bin/bash
# debug variant of the library with clock64() phase counters in the QR kernels (tools/qr_timing.py); tools/variants/ is git-ignored
set -e
cd "$(dirname "$0")/../tnalg_b200/csrc"
mkdir -p ../../tools/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -DTN_QR_TIMING -c qr_householder.cu -o /tmp/qr_timing.o
nvcc -shared -cudart shared -gencode arch=compute_100a,code=sm_100a -o ../../tools/variants/libtnalg_qr_timing.so lib.o chain_gemm.o chain_gemm_tma.o \
    vector_ops.o effh_plan.o lanczos.o jacobi_svd.o comm.o /tmp/qr_timing.o ed_apply.o jacobi_eigh.o expect.o -ldl
