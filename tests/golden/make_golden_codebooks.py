"""Golden code books: the reference's create_{linear,fp8,dynamic,normal}_map for a spread of
parameters (imported from /root/reference as make_golden.py does; pure Python/torch, CPU).

    python tests/golden/make_golden_codebooks.py      ->  tests/golden/reference_codebooks.npz
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402

LINEAR = [(True, 8, True), (True, 8, False), (False, 8, True), (False, 8, False), (True, 4, True), (False, 6, True),
          (True, 7, False), (True, 2, True)]
FP8 = [(True, 5, 2, 8), (True, 4, 3, 8), (False, 4, 4, 8), (True, 2, 1, 4), (True, 3, 2, 6), (False, 3, 2, 5),
       (True, 3, 4, 8), (False, 5, 3, 8)]
DYNAMIC = [(True, 7, 8), (False, 7, 8), (True, 3, 4), (True, 5, 6), (False, 4, 5), (True, 2, 3)]
NORMAL = [(0.9677083, True), (0.9677083, False), (0.95, True), (0.99, False)]


def main():
    bnb = import_reference(with_native=False)
    F = bnb.functional
    out = {}
    for a in LINEAR:
        out["linear_" + "_".join(str(int(v)) for v in a)] = F.create_linear_map(*a).numpy()
    for a in FP8:
        out["fp8_" + "_".join(str(int(v)) for v in a)] = F.create_fp8_map(*a).numpy()
    for a in DYNAMIC:
        out["dynamic_" + "_".join(str(int(v)) for v in a)] = F.create_dynamic_map(*a).numpy()
    for off, extra in NORMAL:
        out[f"normal_{off}_{int(extra)}"] = F.create_normal_map(off, extra).numpy()
    path = HERE / "reference_codebooks.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()
