import sys, json; sys.path.insert(0, '/root/repo')
import torch, bench, m17_sdr_b200 as m
ctx = m.Context(0)
r = bench.wideband_record(ctx, m, torch, 5)
print(json.dumps({k: r[k] for k in ("device_resident", "e2e", "check")}))
