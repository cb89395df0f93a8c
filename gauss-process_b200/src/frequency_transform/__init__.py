# Frequency-domain transformation (FRF on the GPU; the FFT estimator stays the reference's host code)
from src.frequency_transform.transform import estimate_frequency_response  # noqa: F401
