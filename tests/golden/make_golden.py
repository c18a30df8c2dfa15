"""Regenerates tests/golden/ref_alm_golden.npz from the reference's own golden alm files.

Run in the build container only (needs /root/reference): python tests/golden/make_golden.py
The text files are the outputs of python pixell.curvedsky.map2alm printed with %.60g by the reference's
test/data_gen/gen_sht_test_data.py:1-43; they are what test/test_transforms.jl:11-77 compares against.
Only the numbers are kept (float64 binary), keyed by the reference file name.
"""
import os
import numpy as np

SRC = "/root/reference/test/data"
FILES = ["simple_analytic_sht.txt", "simple_analytic_sht_sliced.txt", "simple_box_analytic_sht.txt",
         "simple_analytic_sht_fullalm.txt", "simple_pol_analytic_sht.txt", "test_cls_IQU.txt"]

out = {}
for f in FILES:
    out[f.replace(".txt", "")] = np.loadtxt(os.path.join(SRC, f))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_alm_golden.npz"), **out)
for k, v in out.items():
    print(k, v.shape)

# pixel-area vectors (python pixell's pixsizemap rows; test/test_geometry.jl:287-316 compares pixareamap against them) and the
# reference's FITS I/O fixture (test/test_io.jl:4-14: a 100x100x3 Float64 CAR map), kept byte for byte
import shutil
HERE = os.path.dirname(os.path.abspath(__file__))
np.savez_compressed(os.path.join(HERE, "ref_pixareas.npz"), fullsky=np.loadtxt(os.path.join(SRC, "fullsky_pixareas.dat")),
                    box=np.loadtxt(os.path.join(SRC, "box_pixareas.dat")))
shutil.copyfile(os.path.join(SRC, "test.fits"), os.path.join(HERE, "ref_test.fits"))
os.chmod(os.path.join(HERE, "ref_test.fits"), 0o644)
