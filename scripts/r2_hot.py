"""Top stall lines of one ncu report (source page): python scripts/r2_hot.py gpurun_out/x.ncu-rep [top]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from stall_hotspots import section
print(section(sys.argv[1], sys.argv[1], top=int(sys.argv[2]) if len(sys.argv) > 2 else 24))
