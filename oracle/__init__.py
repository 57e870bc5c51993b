"""TEST INFRASTRUCTURE ONLY - CPU restatements of the reference Kalman filter hot path.

Nothing under oracle/ is product code.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import, call, link or execute it, and only as the checker / the reported CPU
baseline - never as the thing measured or shipped.  The product package optistate_b200/ has no
import of this directory and fails loudly when its CUDA extension is missing.

Parity pin: the reference ships no tests or golden vectors for this path,
so the pins are minted here: oracle/gen_golden.py runs the UNMODIFIED reference class (through
oracle/ref_shim.py, in the build container where /root/reference exists) and commits its outputs
as tests/golden/*.npz; tests/test_oracle.py checks kf_numpy.py and kf_oracle.c against those
fixtures and, when the reference tree is present, against the live reference.
"""
