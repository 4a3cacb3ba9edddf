"""ncu driver: superop -> PTM at n = 3 (pl_pass_a_kernel<3, true>), 16384 matrices (1 GB)."""
import sys, torch
sys.path.insert(0, ".")
import bench_kernels as bk
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
a = bk._rand_c128(torch, (16384, 64, 64), 5)
b = torch.empty_like(a)
for _ in range(3):
    st.superop2pauli_liouville_batch(a, out=b)
torch.cuda.synchronize()
