"""Quick GPU probe: parity report + timing of render fwd/bwd at the BASELINE config-2 shape."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import util_neus as U

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
reps = []
if "--parity" in sys.argv:
    for name, bg, an in [("tiny", "ray", 1.0), ("skiplast", "none", 0.0), ("shipped", "ray", 1.0), ("b2", "white", 1.0)]:
        reps.append(U.run_case_gpu_vs_oracle(name, bg_kind=bg, anneal=an))
        print(json.dumps(reps[-1]))
    json.dump(reps, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)

engine = int(os.environ.get("AVC_ENGINE", "0"))
for case, R in [("b2", 512), ("shipped", 512), ("shipped", 12544)]:
    sdf_kw, col_kw, ren_kw, _ = U.CASES[case]
    sp, cp = U.synth_state(sdf_kw, col_kw, 0)
    sdf, col, var, ren = U.build_product(sdf_kw, col_kw, ren_kw, sp, cp, 0.3, "cuda", engine=engine, chunk=16384)
    o, d, near, far, jit = [t.cuda() for t in U.make_rays(R, 1)]
    bg = torch.rand(R, 1, device="cuda")
    def step():
        out = ren.render(o, d, near, far, background_rgb=bg, cos_anneal_ratio=1.0, jitter=jit)
        loss = out["extra_color_fine"].mean() + 0.1 * out["gradient_error"] + out["weight_sum"].mean()
        loss.backward()
        return out
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    N = 10
    for _ in range(N):
        e0.record()
        out = ren.render(o, d, near, far, background_rgb=bg, cos_anneal_ratio=1.0, jitter=jit)
        loss = out["extra_color_fine"].mean() + 0.1 * out["gradient_error"] + out["weight_sum"].mean()
        e1.record()
        loss.backward()
        e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    print(f"[probe] {case} R={R} engine={engine}: fwd {tf/N:.3f} ms  bwd {tb/N:.3f} ms  total {(tf+tb)/N:.3f} ms", flush=True)
