# Gaussian Process Regression module (kernels, fit/predict, grid search on the GPU)
from src.gpr.kernels import create_kernel, Kernel  # noqa: F401
from src.gpr.gpr_fitting import GaussianProcessRegressor  # noqa: F401
