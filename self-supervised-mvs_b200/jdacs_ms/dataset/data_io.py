"""Mirror of jdacs-ms `dataset/data_io.py` (:15-80): the same PFM reader / writer as the jdacs tree."""
from ...jdacs.datasets.data_io import read_pfm, save_pfm, save_pfm_flipped  # noqa: F401
