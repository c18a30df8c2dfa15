"""CPU oracle for the map2alm / alm2map hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (pixell.jl_b200/) never does; it fails loudly without the CUDA library.
"""
from .oracle import (Oracle, get_oracle, CpuSht, get_cpu_sht, cc_weights, cc_geometry, alm_index, nalm, build)  # noqa: F401
