import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from gemm_bench import bench
v = sys.argv[1] if len(sys.argv) > 1 else "none"
bench(63744, 1024, 256, v, iters=3)
