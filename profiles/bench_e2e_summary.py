"""One-line summary of a bench.py JSON line read from stdin (A/B runs of the host-side paths)."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
g = lambda k: (d.get(k) or {}).get("ms_per_step")
par = d.get("parity") or {}
print(sys.argv[1] if len(sys.argv) > 1 else "", "step %.4f ms" % d["ms_per_step"], "e2e", g("e2e"), "full upload", g("e2e_full_upload"),
      "pageable", g("e2e_pageable"), "call by call", g("e2e_call_by_call"), "parity", par.get("pass"),
      par.get("e2e_theta_identical_to_resident"), par.get("e2e_fir_rmse_rel_err_vs_resident"), "fir_rmse", (d.get("result") or {}).get("e2e_fir_rmse"))
