#!/usr/bin/env python
"""Copy the hot path's CALLERS of the reference, unmodified, into the two git-ignored `_ref` trees.

    python tools/vendor_reference.py [--reference /root/reference]

Runs only where /root/reference exists (the authoring container); `__graft_entry__.build()` calls it.  Nothing is
edited: the files are byte-identical copies, the two target trees are listed in .gitignore (never in history) and
travel to the GPU box with the repo snapshot exactly like the built .so does.

    baseline/_ref/DissimilarDomains/   the "reference checkout" a user of GA-GAN already has.  The product is a
                                       drop-in for the checkout's `torch_utils.ops` + `modulated_conv2d`
                                       (`gagan_b200.install(<checkout>)`); tests and bench.py run the checkout's OWN
                                       training/networks.py, training/loss.py, training/augment.py on top of the
                                       library, and `bench.py --impl reference` times it on the host cores.
    oracle/_ref/DissimilarDomains/     the oracle's private copy: imported under private module names by
                                       oracle/live_ref.py as the CPU `impl='ref'` ground truth of the parity tests.

Only the packages the hot path's callers import are taken (torch_utils, training, dnnlib), plus -- for
row f4, the rosinality adapter -- the five files of SimilarDomains' second StyleGAN2 implementation that hold its hot path
(gan_models/StyleGAN2/{model,nvidia}.py and the torch-native op/ files that op/__init__.py selects); CLI, metrics, dataset
tooling, the rest of SimilarDomains and GA are out of scope and stay where they are.
"""
import os
import sys
import shutil
import hashlib
import argparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKAGES = ('torch_utils', 'training', 'dnnlib')
TARGETS = (os.path.join('baseline', '_ref'), os.path.join('oracle', '_ref'))
SD_FILES = ('gan_models/__init__.py', 'gan_models/StyleGAN2/model.py', 'gan_models/StyleGAN2/nvidia.py', 'gan_models/StyleGAN2/op/__init__.py',
            'gan_models/StyleGAN2/op/upfirdn2d_torch_native.py', 'gan_models/StyleGAN2/op/fused_act_torch_native.py')


def tree_digest(path):
    h = hashlib.sha256()
    for dirpath, dirnames, filenames in sorted(os.walk(path)):
        dirnames[:] = sorted(d for d in dirnames if d != '__pycache__')
        for fn in sorted(filenames):
            if fn.endswith('.pyc'):
                continue
            full = os.path.join(dirpath, fn)
            h.update(os.path.relpath(full, path).encode())
            with open(full, 'rb') as f:
                h.update(f.read())
    return h.hexdigest()


def vendor(reference='/root/reference', verbose=True):
    src_root = os.path.join(reference, 'DissimilarDomains')
    if not os.path.isdir(src_root):
        if verbose:
            print(f'vendor_reference: {src_root} not found, nothing to do (the GPU box uses the trees that travelled with the repo)')
        return False
    for target in TARGETS:
        dst_root = os.path.join(ROOT, target, 'DissimilarDomains')
        for pkg in PACKAGES:
            src, dst = os.path.join(src_root, pkg), os.path.join(dst_root, pkg)
            if os.path.isdir(dst) and tree_digest(src) == tree_digest(dst):
                continue
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        for fn in ('LICENSE.txt',):
            if os.path.isfile(os.path.join(src_root, fn)):
                shutil.copy2(os.path.join(src_root, fn), os.path.join(dst_root, fn))
        with open(os.path.join(dst_root, 'VENDORED.txt'), 'w') as f:
            f.write('Unmodified copy of /root/reference/DissimilarDomains/{torch_utils,training,dnnlib} made by tools/vendor_reference.py.\n'
                    'Git-ignored; not product source.\n')
            for pkg in PACKAGES:
                f.write(f'{pkg} sha256 {tree_digest(os.path.join(dst_root, pkg))}\n')
        sd_src, sd_dst = os.path.join(reference, 'SimilarDomains'), os.path.join(ROOT, target, 'SimilarDomains')
        for rel in SD_FILES:
            src, dst = os.path.join(sd_src, rel), os.path.join(sd_dst, rel)
            if not os.path.isfile(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.isfile(dst) or open(src, 'rb').read() != open(dst, 'rb').read():
                shutil.copy2(src, dst)
        ga_src = os.path.join(reference, 'GA', 'crossover_mutation.py')            # the GA operators (19 lines; the package around
        if os.path.isfile(ga_src) and target.startswith('oracle'):                 # them is un-importable)
            os.makedirs(os.path.join(ROOT, target, 'GA'), exist_ok=True)
            shutil.copy2(ga_src, os.path.join(ROOT, target, 'GA', 'crossover_mutation.py'))
        if verbose:
            print(f'vendor_reference: {dst_root} up to date')
    return True


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    sys.exit(0 if vendor(ap.parse_args().reference) else 1)
