"""TEST INFRASTRUCTURE ONLY.

CPU (numpy/scipy) restatement of the reference's randomized-sketching hot path. It is the
checker for the CUDA product path and the CPU baseline leg of ``bench.py``; nothing under
``parla_b200/`` may import it.  See ``oracle/parla_oracle.py`` for the pinning status.
"""
