"""Stage-by-stage GPU diagnostics, each stage in its own subprocess with a timeout so that a trapped
or hung kernel cannot take the later stages down.  Writes gpurun_out/check.json.

    python tools/gpu_check.py [stage ...]        (default: all stages)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

LAYERS = [
    (256, 256, 64, 1, "reflect", False, 1), (288, 256, 64, 1, "reflect", False, 1), (288, 256, 64, 1, "zeros", False, 1),
    (768, 256, 64, 1, "zeros", False, 1), (64, 128, 256, 2, "zeros", False, 1), (64, 64, 256, 2, "zeros", False, 1),
    (128, 256, 128, 2, "zeros", False, 1), (128, 128, 128, 2, "zeros", False, 1), (256, 128, 64, 2, "zeros", True, 1),
    (128, 64, 128, 2, "zeros", True, 1),
]


def stage_conv(impl, idx):
    import torch
    import torch.nn.functional as F
    import animateportrait_b200 as ap
    Cin, Cout, S, stride, pad_mode, transposed, B = LAYERS[idx]
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(idx)
    x = torch.randn(B, Cin, S, S, generator=g)
    w = torch.randn((Cin, Cout, 3, 3) if transposed else (Cout, Cin, 3, 3), generator=g) * 0.05
    if transposed:
        ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    elif pad_mode == "reflect":
        ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, stride=stride)
    else:
        ref = F.conv2d(x, w, stride=stride, padding=1)
    y, st = ap.conv2d_debug(x.to(dev), w.to(dev), stride=stride, pad=1, pad_mode=pad_mode, transposed=transposed, impl=impl)
    y = y.cpu()
    d = (y - ref).abs()
    rms = ref.pow(2).mean().sqrt().item()
    res = {"layer": LAYERS[idx], "impl": impl, "max_err": d.max().item(), "rms": rms, "rel": d.max().item() / rms,
           "mean_err": d.mean().item(), "stat_sum_err": (st.cpu()[..., 0] - ref.double().sum((2, 3))).abs().max().item()}
    if res["rel"] > 1e-2:
        # where is it wrong?  per-row / per-col / per-channel error maps help to spot layout bugs
        e = d[0]
        res["err_by_channel_first8"] = e.amax((1, 2))[:8].tolist()
        res["err_by_row_first8"] = e.amax((0, 2))[:8].tolist()
        res["err_by_col_first8"] = e.amax((0, 1))[:8].tolist()
        res["err_by_row_last4"] = e.amax((0, 2))[-4:].tolist()
        res["err_by_col_last4"] = e.amax((0, 1))[-4:].tolist()
        res["y_sample"] = y[0, 0, 0, :6].tolist()
        res["ref_sample"] = ref[0, 0, 0, :6].tolist()
        res["frac_bad"] = (d > 1e-2 * rms).float().mean().item()
    return res


def stage_forward(precision, case="c1_line_bias"):
    import torch
    import animateportrait_b200 as ap
    from oracle import netg_oracle as O
    from tests.golden.make_golden import CASES
    onc, B, wseed, bstd, iseed, kind = CASES[case]
    sd = O.make_state_dict(onc, seed=wseed, bias_std=bstd)
    inputs = O.make_inputs(B, seed=iseed, kind=kind)
    dev = torch.device("cuda", 0)
    net = ap.define_G(3, onc, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [0], div=3, disp=3, precision=precision).module
    net.load_state_dict(sd)
    with torch.no_grad():
        y = net(*[t.to(dev) for t in inputs])
    torch.cuda.synchronize()
    taps = {}
    y_ref = O.netg_forward(sd, *inputs, tap=lambda n, v: taps.__setitem__(n, v))
    rep = {}
    for k, ref in taps.items():
        if k == "pre_tanh":
            continue
        got = net.debug_read(k).cpu()
        d = (got - ref).abs()
        rep[k] = {"max": d.max().item(), "mean": d.mean().item(), "ref_absmax": ref.abs().max().item()}
    d = (y.cpu() - y_ref).abs()
    rep["out"] = {"max": d.max().item(), "mean": d.mean().item()}
    rep["launches"] = net.last_launch_count()
    return rep


def stage_env():
    import torch
    p = torch.cuda.get_device_properties(0)
    smi = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total", "--format=csv"],
                         capture_output=True, text=True).stdout
    return {"name": p.name, "sms": p.multi_processor_count, "cc": [p.major, p.minor], "cpus": os.cpu_count(), "smi": smi,
            "torch": torch.__version__}


def run_child(spec, timeout):
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(spec)], capture_output=True,
                           text=True, timeout=timeout, cwd=ROOT)
        out = r.stdout.strip().splitlines()
        res = None
        for line in reversed(out):
            if line.startswith("{"):
                res = json.loads(line)
                break
        if res is None:
            res = {"error": "no result", "rc": r.returncode, "stderr": r.stderr[-1500:], "stdout": r.stdout[-500:]}
    except subprocess.TimeoutExpired:
        res = {"error": f"timeout after {timeout}s"}
    res["_seconds"] = round(time.time() - t0, 1)
    return res


def child(spec):
    kind = spec["kind"]
    try:
        if kind == "env":
            res = stage_env()
        elif kind == "conv":
            res = stage_conv(spec["impl"], spec["idx"])
        elif kind == "forward":
            res = stage_forward(spec["precision"], spec.get("case", "c1_line_bias"))
        else:
            res = {"error": "unknown stage"}
    except Exception as e:  # noqa: BLE001
        import traceback
        res = {"error": repr(e)[:800], "trace": traceback.format_exc()[-1200:]}
    print(json.dumps(res))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(json.loads(sys.argv[2]))
        return
    want = set(sys.argv[1:])
    os.makedirs(OUT, exist_ok=True)
    results = {}

    def go(name, spec, timeout=300):
        if want and not any(name.startswith(w) for w in want):
            return
        results[name] = run_child(spec, timeout)
        print(name, json.dumps(results[name])[:600], flush=True)
        with open(os.path.join(OUT, "check.json"), "w") as f:
            json.dump(results, f, indent=1)

    go("env", {"kind": "env"})
    go("forward_simt", {"kind": "forward", "precision": "fp32_simt"}, 600)
    for i in range(len(LAYERS)):
        go(f"conv_simt_{i}", {"kind": "conv", "impl": "fp32_simt", "idx": i})
    for impl in ("fp32", "bf16"):
        for i in range(len(LAYERS)):
            go(f"conv_{impl}_{i}", {"kind": "conv", "impl": impl, "idx": i}, 120)
    go("forward_fp32", {"kind": "forward", "precision": "fp32"}, 600)
    go("forward_bf16", {"kind": "forward", "precision": "bf16"}, 600)


if __name__ == "__main__":
    main()
