# FIR model identification and validation (hot path: fir_helpers, fir_fitting paper mode, fir_validation; next row:
# kernel_regularized -- the time-domain kernel-regularised FIR marginal likelihood on the device).
# The reference's other modules in this package (fir_legacy, lms_filter, rls_filter, partial_update_lms) are outside
# the hot path; with a reference checkout on src.__path__ they resolve to the reference's own files.
