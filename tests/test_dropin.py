"""The drop-in shims are importable exactly the way the reference's train.py imports its modules
(`from models.mvsnet import MVSNet, mvsnet_loss`, `from losses.unsup_loss import *`, ...), one tree per process."""
import os
import subprocess
import sys

from conftest import ROOT

SNIPPETS = {
    "jdacs": ("from models.mvsnet import MVSNet, mvsnet_loss\nfrom models.module import *\nfrom losses.unsup_loss import *\nfrom models.augmentations import random_image_mask, aug_loss\n"
              "from losses.homography import inverse_warping\nfrom losses.modules import SSIM, depth_smoothness\n"
              # (the reference's own datasets/__init__.py makes `datasets` a package; the overlay adds this one file to it)
              "import importlib.util, os\nsp = importlib.util.spec_from_file_location('data_io', os.path.join(os.environ['PYTHONPATH'], 'datasets', 'data_io.py'))\n"
              "dio = importlib.util.module_from_spec(sp); sp.loader.exec_module(dio); assert callable(dio.read_pfm) and callable(dio.save_pfm)\n"
              "import inspect\nm = MVSNet(refine=False)\nassert list(inspect.signature(m.forward).parameters) == ['imgs', 'proj_matrices', 'depth_values']\n"
              "assert callable(homo_warping) and callable(depth_regression) and UnSupLoss is not None\nprint('ok')"),
    "jdacs-ms": ("from models.network import CVPMVSNet, sL1_loss, MSE_loss\nfrom models.modules import *\nfrom losses.unsup_loss import *\nfrom dataset.data_io import read_pfm, save_pfm\n"
                 "import inspect\nfrom types import SimpleNamespace\nm = CVPMVSNet(SimpleNamespace(nsrc=2, nscale=2, mode='train'))\n"
                 "assert list(inspect.signature(m.forward).parameters) == ['ref_img', 'src_imgs', 'ref_in', 'src_in', 'ref_ex', 'src_ex', 'depth_min', 'depth_max']\n"
                 "assert callable(proj_cost) and callable(calDepthHypo) and callable(homo_warping)\nprint('ok')"),
}


def test_reference_style_imports():
    for tree, code in SNIPPETS.items():
        env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin", tree), SSMVS_B200_ROOT=ROOT)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (tree, r.stderr[-2000:])
