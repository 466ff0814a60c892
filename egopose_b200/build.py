"""Builds libegopose_b200.so (C ABI, include/egopose_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libegopose_b200.so')
SOURCES = ['update_kernels.cu', 'rollout.cu', 'ozaki.cu', 'oz_mlp.cu', 'lstm.cu', 'p2p.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--shared', '-Xptxas', '-v']


def nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.exists(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), 'include', 'egopose_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(src, obj, extra):
    cmd = [nvcc()] + [f for f in NVCC_FLAGS if f != '--shared'] + extra + ['-c', '-o', obj, src]
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, ' '.join(cmd) + '\n' + res.stdout + res.stderr


def build(force=False, verbose=False):
    """one object per .cu (compiled in parallel, rebuilt only when the source or a header is newer), then one link"""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get('EGP_NVCC_EXTRA', '').split()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    tag = os.path.join(objdir, 'flags.txt')
    flags_now = ' '.join(NVCC_FLAGS + extra)
    if not os.path.exists(tag) or open(tag).read() != flags_now:
        force = True
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'egopose_b200.h'))
    hdr_t = max(os.path.getmtime(h) for h in hdrs)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + '.o')
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    log = ''
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for rc, out in ex.map(lambda j: _compile_one(j[0], j[1], extra), jobs):
            log += out
            if rc != 0:
                with open(os.path.join(HERE, 'build.log'), 'w') as f:
                    f.write(log)
                raise RuntimeError('nvcc failed:\n' + out[-4000:])
    cmd = [nvcc(), '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + \
          [os.path.join(objdir, s[:-3] + '.o') for s in SOURCES]
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']
    res = subprocess.run(cmd, capture_output=True, text=True)
    log += ' '.join(cmd) + '\n' + res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(log)
    if verbose:
        print(log)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + log[-4000:])
    open(tag, 'w').write(flags_now)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
