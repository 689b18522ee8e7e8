"""Run one of the reference's CLI scripts (train_transformed_rays.py / eval_transformed_rays.py) UNMODIFIED with the
drop-in `nerf` package of this repo first on sys.path.  Provides minimal stand-ins for `imageio` and
`matplotlib` when those packages are absent (the scripts import them before `nerf`).

    python 4d-facial-avatars_b200/run_reference_script.py <script.py> [script args...]
"""
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def _ensure_imageio():
    try:
        import imageio  # noqa: F401
        return
    except ImportError:
        pass
    import numpy as np
    m = types.ModuleType("imageio")

    def imread(path):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3:
            img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
        return img

    def imwrite(path, img):
        import cv2
        img = np.asarray(img)
        if img.ndim == 3 and img.shape[2] >= 3:
            img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
        cv2.imwrite(path, img)

    m.imread, m.imwrite, m.imsave = imread, imwrite, imwrite
    sys.modules["imageio"] = m


def _ensure_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass

    class _Anything:
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, name):
            return self

    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.use = lambda *a, **k: None
    for name in ("figure", "savefig", "close", "imshow", "show", "plot", "axis", "subplots", "Axes"):
        setattr(plt, name, _Anything())
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    _ensure_imageio()
    _ensure_matplotlib()
    sys.path.insert(0, HERE)  # the drop-in `nerf` wins over the one next to the script
    sys.argv = [script] + sys.argv[2:]
    os.chdir(os.path.dirname(script))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
