"""Test helper: the reference orchestrator's GP path (src/pipeline/gp_pipeline.py:90-156,207-265,269-305),
restated compactly on top of the PUBLIC drop-in surface (src.frequency_transform, src.gpr, src.fir_model),
so that the whole pipeline can be compared with tests/golden/pipeline.npz on a box that has no reference
checkout.  Test infrastructure only."""
from pathlib import Path

import numpy as np
from sklearn.preprocessing import StandardScaler


def run(train_mat, val_mat, out_dir, kernel="matern52", nd=50, noise=1e-6, optimize=False, grid=False,
        max_comb=5000, n_restarts=3, mode="separate", kernel_kwargs=None):
    from src.frequency_transform.transform import estimate_frequency_response
    from src.frequency_transform.frf_estimator import (load_time_u_y, matlab_freq_grid, synchronous_coefficients_trapz,
                                                       frf_cross_power_average)
    from src.gpr.kernels import create_kernel
    from src.gpr.gpr_fitting import GaussianProcessRegressor
    from src.gpr.visualization import plot_gp_results
    from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline

    out_dir = Path(out_dir); out_dir.mkdir(parents=True, exist_ok=True)
    df = estimate_frequency_response(mat_files=[str(train_mat)], method="frf", nd=nd, time_duration=None, n_files=1)
    df.to_csv(out_dir / "unified_frf.csv", index=False)           # data_loader.py:87-88 (CSV round trip)
    import pandas as pd
    frf = pd.read_csv(out_dir / "unified_frf.csv")
    omega = frf["omega_rad_s"].values
    G = frf["ReG"].values + 1j * frf["ImG"].values
    Xs = StandardScaler(); Xn = Xs.fit_transform(np.log10(omega.reshape(-1, 1)))
    ysr = StandardScaler(); yr = ysr.fit_transform(G.real.reshape(-1, 1)).ravel()
    ysi = StandardScaler(); yi = ysi.fit_transform(G.imag.reshape(-1, 1)).ravel()
    vX = vyr = vyi = None
    if grid and val_mat is not None:                               # gp_pipeline.py:90-108, data_loader.py:137-203
        t, u, y = load_time_u_y(Path(val_mat).resolve(), y_col=0)
        _, wv = matlab_freq_grid(150, f_low_log10=-1.0, f_up_log10=2.3)
        Gv = frf_cross_power_average([synchronous_coefficients_trapz(t, u, wv, 0.0, True)],
                                     [synchronous_coefficients_trapz(t, y, wv, 0.0, True)])
        vX = Xs.transform(np.log10(wv).reshape(-1, 1)); vyr, vyi = np.real(Gv), np.imag(Gv)
    if mode == "polar":                                            # gp_pipeline.py:159-193
        return _run_polar(omega, G, Xs, Xn, frf, kernel, noise, optimize, grid, max_comb, n_restarts, val_mat, out_dir)
    fitted = []
    for target, scaler, vy in ((yr, ysr, vyr), (yi, ysi, vyi)):
        gp = GaussianProcessRegressor(kernel=create_kernel(kernel, **(kernel_kwargs or {})), noise_variance=noise)   # gp_pipeline.py:282-286
        gp.fit(Xn, target, optimize=optimize, n_restarts=n_restarts, use_grid_search=grid,
               max_grid_combinations=max_comb, validation_X=vX, validation_y=vy, y_scaler=scaler)
        pred, std = gp.predict(Xn, return_std=True)
        fitted.append((gp, scaler.inverse_transform(pred.reshape(-1, 1)).ravel(), std * scaler.scale_))
    (gp_r, pr, sr), (gp_i, pi_, si) = fitted
    res = {"real": plot_gp_results(omega, G.real, pr, sr, "", None), "imag": plot_gp_results(omega, G.imag, pi_, si, "", None),
           "kernel_params": {"real": gp_r.kernel.get_params().tolist(), "imag": gp_i.kernel.get_params().tolist()},
           "frf": frf.to_numpy()}

    def predict_at(omega_new):                                     # gp_pipeline.py:207-226
        X = omega_new.reshape(-1, 1).copy()
        omin = np.min(omega[omega > 0])
        X[X.ravel() <= omin * 0.1] = omin
        Xq = Xs.transform(np.log10(X))
        a = ysr.inverse_transform(gp_r.predict(Xq).reshape(-1, 1)).ravel()
        b = ysi.inverse_transform(gp_i.predict(Xq).reshape(-1, 1)).ravel()
        return a + 1j * b

    res["fir_extraction"] = gp_to_fir_direct_pipeline(omega=omega, G=pr + 1j * pi_, gp_predict_func=predict_at,
                                                      mat_file=Path(val_mat) if val_mat else None,
                                                      output_dir=out_dir, fir_length=1024)
    return res


def _run_polar(omega, G, Xs, Xn, frf, kernel, noise, optimize, grid, max_comb, n_restarts, val_mat, out_dir):
    """gp_mode='polar': GPs on log|G| and the unwrapped phase; the search gets no validation data, so its metric
    is the NLL (gp_pipeline.py:159-193), and the predictor recombines exp(a) * exp(jb) (:207-226)."""
    from src.gpr.kernels import create_kernel
    from src.gpr.gpr_fitting import GaussianProcessRegressor
    from src.gpr.visualization import plot_gp_results
    from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline

    mag = np.abs(G); phase = np.unwrap(np.angle(G))
    ms = StandardScaler(); ym = ms.fit_transform(np.log(mag).reshape(-1, 1)).ravel()
    ps = StandardScaler(); yp = ps.fit_transform(phase.reshape(-1, 1)).ravel()
    fitted = []
    for target, scaler in ((ym, ms), (yp, ps)):
        gp = GaussianProcessRegressor(kernel=create_kernel(kernel), noise_variance=noise)
        gp.fit(Xn, target, optimize=optimize, n_restarts=n_restarts, use_grid_search=grid, max_grid_combinations=max_comb)
        pred, std = gp.predict(Xn, return_std=True)
        fitted.append((gp, scaler.inverse_transform(pred.reshape(-1, 1)).ravel(), std * scaler.scale_))
    (gp_m, pm, _), (gp_p, pp, sp) = fitted
    mag_pred = np.exp(pm)
    res = {"magnitude": plot_gp_results(omega, mag, mag_pred, None, "", None),
           "phase": plot_gp_results(omega, phase, pp, sp, "", None),
           "kernel_params": {"magnitude": gp_m.kernel.get_params().tolist(), "phase": gp_p.kernel.get_params().tolist()},
           "frf": frf.to_numpy()}

    def predict_at(omega_new):
        X = omega_new.reshape(-1, 1).copy()
        omin = np.min(omega[omega > 0])
        X[X.ravel() <= omin * 0.1] = omin
        Xq = Xs.transform(np.log10(X))
        a = ms.inverse_transform(gp_m.predict(Xq).reshape(-1, 1)).ravel()
        b = ps.inverse_transform(gp_p.predict(Xq).reshape(-1, 1)).ravel()
        return np.exp(a) * np.exp(1j * b)

    res["fir_extraction"] = gp_to_fir_direct_pipeline(omega=omega, G=mag_pred * np.exp(1j * pp), gp_predict_func=predict_at,
                                                      mat_file=Path(val_mat) if val_mat else None,
                                                      output_dir=out_dir, fir_length=1024)
    return res
