"""LLFF loading is outside the NeRFace render path; the name is kept importable because the
reference's scripts import it (train_transformed_rays.py:17-21)."""


def load_llff_data(*args, **kwargs):
    raise NotImplementedError("LLFF datasets are out of scope of the B200 NeRFace render path")
