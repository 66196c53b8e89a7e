"""one config of bench_configs.py on its own (for ncu launch lists):  python scripts/config_probe.py 3"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs
from parth_b200 import api
which = sys.argv[1] if len(sys.argv) > 1 else "3"
kind = "reference" if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libparth_ref.so")) else "port"
print(json.dumps(bench_configs.run_all(api, kind, 6545.6, (which,))))
