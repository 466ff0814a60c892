"""The reference's UNMODIFIED ego_pose/ego_mimic.py / ego_forecast.py executed through egopose_b200/compat (CPU part): the
scripts must get past argument parsing, Config, seeding, Logger / create_logger and reach the first CUDA requirement
(HumanoidEnv -> libegopose_b200 model creation), which fails loudly with EgpError on a machine without a GPU.  Skipped when
/root/reference is absent (the GPU box); the GPU run of the same script is logged under profiles/."""
import os
import subprocess
import sys
import tempfile

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
@pytest.mark.parametrize('script', ['ego_pose/ego_mimic.py', 'ego_pose/ego_forecast.py'])
def test_unmodified_script_reaches_the_cuda_boundary(script):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present: the full run is exercised by tools/run_reference_script.py')
    work = tempfile.mkdtemp(prefix='egp_dropin_test_')
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import run_reference_script as r\n"
            "task = 'egoforecast' if 'forecast' in %r else 'egomimic'\n"
            "r.prepare(%r, task, 'dropin_01', 2, 4000, 1, write_data=False)\n"
            "r.run(%r, %r, 'dropin_01', [])\n") % (ROOT, os.path.join(ROOT, 'tools'), script, work, os.path.join(REF, script), work)
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    err = res.stderr
    assert res.returncode != 0
    assert 'NameError' not in err and 'ImportError' not in err and 'ModuleNotFoundError' not in err, err[-3000:]
    assert 'EgpError' in err and ('CUDA device is required' in err or 'libegopose_b200.so is missing' in err), err[-3000:]
    # the script got as far as creating its loggers in the run directory
    assert os.path.isdir(os.path.join(work, 'results'))
