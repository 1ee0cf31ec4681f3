"""Markdown summary of the committed bench lines (reads profiles/r02/bench_n*_final.json).
  python profiles/r02/make_table.py           prints the table
  python profiles/r02/make_table.py --write   also rewrites the block between the `results:begin/end` markers of DESIGN.md and README.md"""
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(n):
    p = os.path.join(HERE, "bench_n%d_final.json" % n)
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


def main():
    rows = []
    base = None
    for n in (1, 2, 4, 8):
        d = load(n)
        if d is None:
            continue
        if n == 1:
            base = d["value"]
        k = d["kernels"]
        fwd = next(v for kk, v in k.items() if kk.startswith("attn_fwd")); bwd = next(v for kk, v in k.items() if kk.startswith("attn_bwd"))
        gemm = sum(v["ms_per_launch"] * v["launches_per_step"] for kk, v in k.items() if kk.startswith("umma"))
        side = d.get("side_legs", {})
        gb = next((v for kk, v in side.items() if kk.startswith("global_batch")), None)
        nv = d.get("nvlink")
        rows.append("| %d | %.1f M | %.3f | %.2f | %.1f M (%.2f) | %.1f M | %.3f / %.3f | %.3f | %s | %s | %s | %s |" % (
            n, d["value"] / 1e6, d["ms_per_step"], d["value"] / (n * base) if base else 1.0, d["e2e"]["value"] / 1e6, d["e2e"]["value"] / d["value"],
            d["e2e_detail"]["idx"]["value"] / 1e6, fwd["ms_per_launch"], bwd["ms_per_launch"], gemm,
            "%.0f / %.0f GB/s" % (nv["gather_in_GBs"], nv["red_add_out_GBs"]) if nv else "pair %.2f of HBM" % d["roofline"]["fused_pair"]["frac"],
            "%.1f M" % (side["zipf_ids"]["value"] / 1e6) if "zipf_ids" in side else "–",
            "%.1f M (%.3f ms)" % (gb["value"] / 1e6, gb["ms_per_step"]) if gb else "–",
            "ok" if d.get("parity", {}) and d["parity"].get("ok") else ("–" if n == 1 else "?")))
    print("| GPUs | samples/s (`value`) | ms/step | vs N x 1-GPU | e2e key-fed (of value) | e2e id-fed | attn fwd / bwd ms | six GEMMs ms | NVLink in / out (N>1) | Zipf ids | global batch 65 536 | parity |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    d = load(1)
    if d and "cpu_baseline" in d:
        c = d["cpu_baseline"]
        print("\nCPU arm (`cpu_fast.c`, %d threads — the fastest count on the box's %s usable CPUs): %.0f samples/s on the same 65 536-sample step; 1 thread: %.0f samples/s."
              % (c["cores"], c.get("usable_threads", "?"), c["value"], c.get("one_thread_value", 0)))


if __name__ == "__main__":
    if "--write" in sys.argv:
        buf = io.StringIO(); old = sys.stdout; sys.stdout = buf
        main()
        sys.stdout = old
        block = "<!-- results:begin (profiles/r02/make_table.py --write) -->\n" + buf.getvalue().strip() + "\n<!-- results:end -->"
        root = os.path.dirname(os.path.dirname(HERE))
        for name in ("DESIGN.md", "README.md"):
            path = os.path.join(root, name)
            txt = open(path).read()
            a, b = txt.index("<!-- results:begin"), txt.index("<!-- results:end -->") + len("<!-- results:end -->")
            open(path, "w").write(txt[:a] + block + txt[b:])
        print(block)
    else:
        main()
