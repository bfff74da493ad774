import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(f'{d["value"]:.1f} pairs/s  {d["ms_per_step"]:.1f} ms  clk {d["clocks"]["sm_mhz"]}  fwd {r["fwd_ms"]:.1f}  bwd_fused {r["bwd_fused_kernel_ms"]:.1f}  gemms {r["bwd_gemms_ms"]:.1f}  frac {r["frac"]:.4f}')
