"""A/B of the FBPINN windowed bank forward (K = 64 nets 1-[64]x3-1, tanh, u / u_x / u_xx): activation jets on packed FP32
pairs (default) or scalars (AX_TC_BANK_PACKED=0), one block barrier per net (default) or two (AX_TC_BANK_ONEBAR=0): time
per launch at 2^17 and 2^20 points and the largest difference between the outputs of the variants.  Each variant runs in
its own process (the switches are read once)."""
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
        torch.save(out[..., :n].cpu(), f"/tmp/bank_ab_{os.environ.get('AX_BANK_AB_TAG', 'x')}_{n}.pt")
    print(json.dumps(rec))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    import torch

    res = {}
    variants = [("scalar_two_barriers", "0", "0"), ("packed_two_barriers", "1", "0"), ("packed_one_barrier", "1", "1")]
    for rnd in range(2):
        for tag, packed, onebar in variants:
            env = dict(os.environ, AX_TC_BANK_PACKED=packed, AX_TC_BANK_ONEBAR=onebar, AX_BANK_AB_TAG=tag)
            out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            if not line:
                print(out.stdout[-2000:], out.stderr[-2000:])
                sys.exit(1)
            res.setdefault(tag, []).append({k: v["ms"] for k, v in json.loads(line[-1]).items()})
    diff = {}
    for n in (1 << 17, 1 << 20):
        a = torch.load(f"/tmp/bank_ab_scalar_two_barriers_{n}.pt").double()
        for tag, _, _ in variants[1:]:
            b = torch.load(f"/tmp/bank_ab_{tag}_{n}.pt").double()
            diff[f"{tag}_{n}"] = float((a - b).abs().max() / a.abs().max())
    res["max_abs_difference_from_scalar_over_largest_entry"] = diff
    print(json.dumps(res))


if __name__ == "__main__":
    main()
