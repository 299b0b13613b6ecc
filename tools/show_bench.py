"""Pretty-print the JSON line of bench.py (stdin or file)."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` closes the pipe early
src = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
line = [l for l in src.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print("value %.1f %s | %.2f ms/step | e2e %.1f | launches/step %s | clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"), d.get("clocks")))
for k, v in sorted(d.get("kernel_families", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print("  %-20s %8.3f ms  share %.3f  n=%-3d %s %s" % (k, v["ms_per_step"], v["share"], v["launches_per_step"],
          ("%.1f TF/s" % v["tflops"]) if "tflops" in v else "", ("%.0f GB/s" % v["gbs"]) if "gbs" in v else ""))
print("roofline", d.get("roofline"))
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
