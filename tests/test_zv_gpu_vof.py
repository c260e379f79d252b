"""tool_vof (row f4): VofCloud(dictfile, mesh).writeVOF — the reference's phase-field initialiser (tool_vof/main.cpp,
tool_vof/solidcloud.cpp:141-173) through the host façade and sdfibm_volume_fraction, held to G1, the field the reference's own
tool wrote for tool_vof/example (alpha.water: 12 solids + the planes{} block of tool_vof/example/solidDict:134-148)."""
import os

import numpy as np
import pytest

from sdfibm_b200 import cases, hostapi
from sdfibm_b200.mesh import Mesh

pytestmark = pytest.mark.gpu

# tool_vof/example/solidDict restated (shapes :16-60, solids :62-132, planes :134-148)
SHAPES = {
    "circle1": dict(type="Circle", radius=0.3),
    "tail1": dict(type="Circle_Tail", radius=0.3, ratio=1, thickness=0.1),
    "ellipse1": dict(type="Ellipse", radiusa=0.3, radiusb=0.2),
    "rect1": dict(type="Rectangle", radiusa=0.3, radiusb=0.2),
    "plane1": dict(type="Plane"),
}
SOLIDS = [("tail1", (0.5, 1.5, 0), -45), ("tail1", (0.5, 2.5, 0), 0), ("tail1", (0.5, 3.5, 0), 45),
          ("rect1", (1.5, 1.5, 0), 0), ("rect1", (1.5, 2.5, 0), 30), ("rect1", (1.5, 3.5, 0), 60),
          ("ellipse1", (2.5, 1.5, 0), 0), ("ellipse1", (2.5, 2.5, 0), 60), ("ellipse1", (2.5, 3.5, 0), 120),
          ("circle1", (3.5, 1.0, 0), 0), ("circle1", (3.5, 2.5, 0), 0), ("circle1", (3.5, 3.5, 0), 0)]
PLANES = [("plane1", (0, 0, 0), 15), ("plane1", (0, 0, 0), -90)]


def test_write_vof_reproduces_the_shipped_alpha_water(tmp_path, m1_points, g1_alpha):
    mesh = Mesh.hex_block_with_points((200, 200, 1), m1_points)
    body = lambda spec: [dict(shp_name=n, pos=p, euler=(0, 0, e)) for n, p, e in spec]
    path = hostapi.write_vof_dict(os.path.join(str(tmp_path), "solidDict"), True, SHAPES, body(SOLIDS), body(PLANES))
    alpha, total, n_solids, n_planes = hostapi.write_vof(path, str(tmp_path), mesh, "alpha.water")
    assert (n_solids, n_planes) == (12, 2)
    assert np.abs(alpha - g1_alpha).max() <= 1e-15
    assert np.array_equal(alpha > 0, g1_alpha > 0) and (alpha > 0).sum() == 13862
    assert abs(total - 5.201762384934972) < 1e-11              # "total volume =" of the tool's report (:169); summation order differs
    # the written field file holds the same numbers
    txt = open(os.path.join(str(tmp_path), "0_alpha.water")).read().split("(\n", 1)[1].rsplit(")", 1)[0]
    assert np.array_equal(np.array(txt.split(), dtype=float), alpha)


def test_volume_fraction_entry_leaves_the_coupling_state_alone():
    from oracle.oracle_py import Oracle
    from sdfibm_b200.context import Context

    case = cases.case_mixed3d(n=20, n_solids=12)
    ctx = Context(0, cell_slots=8)
    ctx.set_mesh(case["mesh"], False)
    ctx.set_shapes(case["shapes"])
    first = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    alpha, total = ctx.volume_fraction(case["solids"][:7])
    o = Oracle(case["mesh"], False)
    ref = o.interact(case["shapes"], case["solids"][:7], 0 * case["U"], 1.0, 1.0)
    assert np.abs(alpha - ref["As"]).max() <= 1e-12 and abs(total - float((ref["As"] * case["mesh"].V).sum())) <= 1e-10 * total
    # fixInternal still sees the Ct of the interact before
    assert np.array_equal(ctx.fix_internal(case["solids"], case["U"]), o.fix_internal(case["shapes"], case["solids"], first["Ct"], case["U"]))
    ctx.close()
