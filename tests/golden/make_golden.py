"""Generate the committed golden fixtures from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  m1_points.npz M1  tool_vof/example/constant/polyMesh  (real blockMesh output, 200x200x1; points only —
                    faces/owner/neighbour are asserted equal to Mesh.hex_block's and regenerated)
  g1_alpha.npz  G1  tool_vof/example/0/alpha.water      (clipped sum of As of 14 solids, HARD golden)
  m2_points.npz M2  examples/falling_ellipse/constant/polyMesh (100x200x1), same treatment
  g2_As.npz     G2  examples/flow_past_cylinder/re200/0/As, central uniform block only
                    (cells 20800 + i + 120 j of the 9-block mesh; every non-zero value lives there)
Data files only — no reference source code is copied.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfibm_b200 import foam_io  # noqa: E402

REF = os.environ.get("SDFIBM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def check_topology(pm, n):
    """The shipped faces/owner/neighbour equal the blockMesh-numbered generator's, so only the point
    coordinates (which carry blockMesh's rounding) are stored."""
    from sdfibm_b200.mesh import Mesh

    t = Mesh.hex_block(n)
    assert np.array_equal(pm["face_off"], t.fp_off) and np.array_equal(pm["face_pts"], t.fp)
    assert np.array_equal(pm["owner"], t.owner) and np.array_equal(pm["neighbour"], t.neighbour)


def main():
    m1 = foam_io.read_polymesh(os.path.join(REF, "tool_vof/example"))
    check_topology(m1, (200, 200, 1))
    np.savez_compressed(os.path.join(OUT, "m1_points.npz"), points=m1["points"], n=np.array([200, 200, 1]))
    g1 = foam_io.read_scalar_field(os.path.join(REF, "tool_vof/example/0/alpha.water"))
    np.savez_compressed(os.path.join(OUT, "g1_alpha.npz"), alpha=g1)
    m2 = foam_io.read_polymesh(os.path.join(REF, "examples/falling_ellipse"))
    check_topology(m2, (100, 200, 1))
    np.savez_compressed(os.path.join(OUT, "m2_points.npz"), points=m2["points"], n=np.array([100, 200, 1]))
    g2 = foam_io.read_scalar_field(os.path.join(REF, "examples/flow_past_cylinder/re200/0/As"))
    nz = np.nonzero(g2)[0]
    assert nz.min() >= 20800 and nz.max() < 20800 + 14400, (nz.min(), nz.max())
    np.savez_compressed(os.path.join(OUT, "g2_As.npz"), As_central=g2[20800:20800 + 14400], n_total=len(g2),
                        n_nonzero=len(nz))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
