"""A/B of the FBPINN windowed bank forward (K = 64 nets 1-[64]x3-1, tanh, u / u_x / u_xx) with the activation jets
evaluated on packed FP32 pairs (default) or scalars (AX_TC_BANK_PACKED=0): time per launch at 2^17 and 2^20 points and
the largest difference between the two outputs.  Each variant runs in its own process (the switch is read once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch

    from ai4xde_b200 import _lib
    from ai4xde_b200.jet import JetSpec, to_soa

    K = 64
    dev = torch.device("cuda:0")
    js = JetSpec((0,), ((0, 0),))
    widths = [1, 64, 64, 64, 1]
    net = _lib.make_net_desc(widths, "tanh", n_nets=K)
    torch.manual_seed(0)
    params = (torch.randn((K, _lib.param_count(_lib.make_net_desc(widths, "tanh")))) * 0.2).to(dev)
    edges = torch.linspace(0, 1, K + 1)
    lo, hi = edges[:-1].contiguous().to(dev), edges[1:].contiguous().to(dev)
    lst = torch.arange(K, dtype=torch.int32, device=dev)
    rec = {}
    for n in (1 << 17, 1 << 20):
        xs = to_soa(torch.rand((n, 1), generator=torch.Generator().manual_seed(1)).to(dev))
        for _ in range(2):
            out = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, lo, hi, lst, engine="tf32")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, lo, hi, lst, engine="tf32")
        e1.record()
        torch.cuda.synchronize()
        rec[str(n)] = {"ms": e0.elapsed_time(e1) / 5, "checksum": [float(v) for v in out[..., :n].double().abs().sum(dim=-1).reshape(-1)]}
        torch.save(out[..., :n].cpu(), f"/tmp/bank_ab_{os.environ.get('AX_TC_BANK_PACKED', '1')}_{n}.pt")
    print(json.dumps(rec))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    import torch

    res = {}
    for packed in ("0", "1", "0", "1"):
        env = dict(os.environ, AX_TC_BANK_PACKED=packed)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print(out.stdout[-2000:], out.stderr[-2000:])
            sys.exit(1)
        res.setdefault("packed" if packed == "1" else "scalar", []).append({k: v["ms"] for k, v in json.loads(line[-1]).items()})
    diff = {}
    for n in (1 << 17, 1 << 20):
        a, b = torch.load(f"/tmp/bank_ab_0_{n}.pt").double(), torch.load(f"/tmp/bank_ab_1_{n}.pt").double()
        diff[str(n)] = float((a - b).abs().max() / a.abs().max())
    res["max_abs_difference_over_largest_entry"] = diff
    print(json.dumps(res))


if __name__ == "__main__":
    main()
