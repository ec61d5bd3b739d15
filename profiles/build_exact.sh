#!/bin/bash
# Second build of the library with the two deliberate arithmetic deviations of the stage kernel switched off:
#   -DHN_EXACT_DIV    IEEE division instead of rcp.approx + 2 Newton steps (stage_pair.cuh pr_rcp)
#   -DHN_NO_GZ_FLUSH  grad(z_bot) at the quadrature points kept as computed (no flush of derivative noise to zero)
# tests/test_gpu_acceptance.py loads it (when present) next to the default build and records the difference per field in
# profiles/parity_r2.json.  Not part of build(): it is a measurement aid.
cd "$(dirname "$0")/.." && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -diag-suppress 177,550 -DHN_EXACT_DIV -DHN_NO_GZ_FLUSH -o h-numo_b200/libhnumo_b200_exact.so h-numo_b200/csrc/hnumo_b200.cu -lpthread -ldl 2>&1 | grep -i "error" ; ls -la h-numo_b200/libhnumo_b200_exact.so
