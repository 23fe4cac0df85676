"""Teacher-forced per-stage parity table (tests/_taps.py) -> stdout / profiles: python tools/tap_parity.py [config] [batch]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _taps import eager_self_noise, teacher_forced_errors, teacher_forced_op_errors  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vit_b16"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ops_res = teacher_forced_op_errors(name, batch, blocks=None if name != "vit_b16" else (0, 6, 11))
print(f"# teacher-forced PER-OP parity vs eager bf16 autocast, {name}, batch {batch} (relative L2; every op fed eager's own input)")
for k, v in ops_res.items():
    print(f"{k:52s} {v:.3e}")
print(f"# worst op: {max(ops_res.values()):.3e}")
out, gerr = teacher_forced_errors(name, batch)
print(f"# teacher-forced stage parity vs eager bf16 autocast, {name}, batch {batch} (relative L2)")
for k, v in out.items():
    print(f"{k:48s} {v:.3e}")
print("# parameter gradients (stage fed eager's upstream gradient)")
groups = {}
for k, v in gerr.items():
    short = k.split(".", 3)[-1] if k.startswith("encoder.mixing_blocks.") else k
    groups.setdefault(short, []).append(v)
for k, vs in groups.items():
    print(f"{k:48s} max {max(vs):.3e}  median {sorted(vs)[len(vs) // 2]:.3e}  (n={len(vs)})")
noise = eager_self_noise(name, min(batch, 64))
print(f"# context: eager vs ITSELF end to end (flash vs math SDPA backend, batch {min(batch, 64)}): " + ", ".join(f"{k} {v:.3e}" for k, v in noise.items()))
