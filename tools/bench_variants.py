"""Build kernel variants (-D tunables of csrc/*.cu) HERE (no GPU needed), then measure them all in one GPU session:

    python tools/bench_variants.py build                       # on the CPU box: lib/libgh_raster_<name>.so per variant
    python tools/bench_variants.py run [--strands 5000] ...    # on the GPU box: tools/stage_times.py per variant, one table

The variant table lives in VARIANTS below; the product library (default tunables) is always measured as 'product'."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = json.load(open(os.path.join(ROOT, "tools", "variants.json"))) if os.path.isfile(os.path.join(ROOT, "tools", "variants.json")) else {}

def main():
    from gaussianhaircut_b200 import build as ghb
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "build":
        ghb.build(verbose=False)
        for name, defs in VARIANTS.items():
            print("building", name, defs, flush=True)
            ghb.build(verbose=False, extra_flags=[f"-D{d}" for d in defs], variant=name)
        return
    extra = sys.argv[2:]
    libs = [("product", ghb.LIB_PATH)] + [(n, os.path.join(ghb.LIB_DIR, f"libgh_raster_{n}.so")) for n in VARIANTS]
    rows = []
    for name, path in libs:
        if not os.path.isfile(path):
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stage_times.py"), path] + extra, capture_output=True, text=True)
        try:
            r = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            print(name, "FAILED", out.stderr[-600:]); continue
        r["variant"] = name; r["defines"] = VARIANTS.get(name, [])
        rows.append(r)
        st = r["stages_us"]
        print(f"{name:28s} {r['ms_per_step']:.4f} ms/step  fwd {st.get('blend_forward', 0):6.1f}  bwd {st.get('blend_backward', 0):6.1f}  "
              f"pre {st.get('preprocess', 0):5.1f} scan {st.get('tile_scan', 0):5.1f} emit {st.get('emit', 0):5.1f} pbwd {st.get('preprocess_backward', 0):5.1f}", flush=True)
    print(json.dumps(rows))

if __name__ == "__main__":
    main()
