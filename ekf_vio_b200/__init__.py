"""ekf_vio_b200 — B200-native (sm_100a) EKF predict/update and pyramidal KLT tracker of
k-sheridan/ekf_vio behind a C ABI (include/ekfvio_c.h).  Python here is plumbing only."""
__all__ = ["capi", "workload"]
