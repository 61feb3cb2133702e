"""Development probe: forward TFLOP/s for one FA_FWD_EMU setting (set in the environment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_probe import bench
import flash_attention_metal_b200 as fa
print("FA_FWD_EMU =", os.environ.get("FA_FWD_EMU"))
bench(1, 16, 16384, 128, True)
bench(1, 16, 16384, 128, False)
bench(8, 12, 4096, 64, True)
bench(16, 8, 2048, 64, False)
