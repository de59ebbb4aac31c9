"""Generates tests/golden/fixture_golden.npz FROM THE UNMODIFIED REFERENCE (oracle/_ref/libdjbref.so) at the sizes the
reference itself uses, and unpacks the two measured fixtures the reference ships (mitsuba/dj_matpreview.zip) into
tests/_fixtures/ (git-ignored: MERL data is licensed and is not committed; the directory travels to the GPU box with the
repo snapshot, like the built libraries).  Committed: only the reference's OUTPUTS (fitted tables and parameters).

    python tests/golden/make_fixture_golden.py          # needs /root/reference; about a minute (two 90 x 90 fits at 10 s each)

Contents:
  merl_fixture/...   djb::tabular(merl("blue-metallic-paint.binary"), 90) + both fit_*_parameters  (examples/merl_params.cpp)
  utia_fixture/...   djb::tabular_anisotropic(utia("m064_fabric099.bin"), 90, 90) + both 5-parameter fits (mitsuba/dj_brdf.cpp:238)
  utia12_90x90/...   the same 90 x 90 fit of the seeded synthetic UTIA table tests/cases.random_utia_table(12): needs no fixture
"""
import sys
import zipfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import api  # noqa: E402
from tests import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
FIX = ROOT / "tests" / "_fixtures"


def unpack_fixtures():
    z = api.REF_ROOT / "mitsuba" / "dj_matpreview.zip"
    FIX.mkdir(exist_ok=True)
    with zipfile.ZipFile(z) as zf:
        for n in zf.namelist():
            if n.endswith("blue-metallic-paint.binary") or n.endswith("m064_fabric099.bin"):
                (FIX / Path(n).name).write_bytes(zf.read(n))
    return FIX / "blue-metallic-paint.binary", FIX / "m064_fabric099.bin"


def main():
    ref = api.RefOracle()
    merl_path, utia_path = unpack_fixtures()
    merl = np.fromfile(merl_path, dtype=np.float64, offset=12)
    utia = np.fromfile(utia_path, dtype=np.float64)
    g = {"merl_fixture/sha256": np.array([cases.sha(merl)]), "utia_fixture/sha256": np.array([cases.sha(utia)])}
    for k, v in ref.fit_tabular(api.Source.merl(merl), 90).items():
        g[f"merl_fixture/{k}"] = v
    for k, v in ref.fit_tabular_anisotropic(api.Source.utia(utia), 90, 90).items():
        g[f"utia_fixture/{k}"] = v
    for k, v in ref.fit_tabular_anisotropic(api.Source.utia(cases.random_utia_table(12)), 90, 90).items():
        g[f"utia12_90x90/{k}"] = v
    np.savez_compressed(OUT / "fixture_golden.npz", **g)
    print("merl fixture alpha (beckmann, ggx):", g["merl_fixture/alpha"])
    print("utia fixture beckmann:", g["utia_fixture/beckmann"], "ggx:", g["utia_fixture/ggx"])
    print((OUT / "fixture_golden.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
