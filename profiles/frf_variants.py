"""Timing experiments on b200id_frf_project_uniform: every library found under gauss-process_b200/lib/ (the default
build plus any exp_*.so variants built with other -D flags) runs the Goertzel form (1) and the
bin-adaptive form (2: DMMA main term + rounding term for the listed bins) on 10^7 samples at N_d = 100 (multisine)
and 150 (square wave); form 2 also with every bin listed, and the number of bins its decision listed.

    python profiles/frf_variants.py            # on the GPU box; one subprocess per library
"""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "gauss-process_b200" / "lib"


def child():
    sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
    import numpy as np
    import torch
    from b200id import _lib, frf, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    lib = _lib.load()
    n = 10_000_000
    out = {}
    for nd in (100, 150):
        rec = np.load(os.environ["FRF_VARIANT_DATA"] + f"_{nd}.npy")
        t, u, y = rec[0], rec[1], rec[2]
        _, w = synth.log_grid(nd)
        t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
        span = float(t[-1] - t[0])
        t0, dt = float(t[0]), span / (n - 1)
        sums = frf.moments_device(u_d, y_d, (0, n))
        for which in (1, 2, 3):
            lib.b200id_frf_select_recurrence(C.c_int(min(which, 2)))
            lib.b200id_frf_adaptive_force(C.c_int(1 if which == 3 else 0))
            for _ in range(3):
                p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
            e1.record()
            torch.cuda.synchronize()
            out[f"nd{nd}_form{which}_ms"] = round(e0.elapsed_time(e1) / 10, 4)
            out[f"nd{nd}_form{which}_checksum"] = float(p.abs().sum())
            if which >= 2:
                out[f"nd{nd}_form{which}_bins_listed"] = frf.adaptive_diag(nd, dev)[0]
                out[f"nd{nd}_form{which}_vs_form1_rel_max"] = float(((p - ref).abs() / ref.abs()).max())
            else:
                ref = p.clone()
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("FRF_VARIANT_CHILD"):
        child()
    else:
        res = {}
        sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
        import numpy as np
        from b200id import synth
        data = "/dev/shm/frf_variant_rec" if os.path.isdir("/dev/shm") else "/tmp/frf_variant_rec"
        for nd in (100, 150):                                   # the records are synthesised once, outside every child
            np.save(data + f"_{nd}.npy", np.stack(synth.make_record(10_000_000, 100, "multisine" if nd == 100 else "square", seed=1)))
        for lib in sorted(LIBDIR.glob("*.so")):
            env = dict(os.environ, FRF_VARIANT_CHILD="1", B200ID_LIB=str(lib), FRF_VARIANT_DATA=data)
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
            last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
            try:
                res[lib.name] = json.loads(last)
            except Exception:
                res[lib.name] = {"error": last}
        print(json.dumps(res, indent=1))
