#!/bin/bash
# Install the unmodified reference (pure Python) into baseline/_ref/ — build container only.
#   pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails twice offline: (1) the build writes egg-info into the read-only source tree -> install from a copy
# under /tmp; (2) setup.py's `setup_requires=['sphinx>=2.1.2']` (documentation tooling) cannot be fetched
# without a network -> that ONE line of the build script is dropped in the /tmp copy (the package sources
# are untouched) and dependency resolution is skipped with --no-deps (torch etc. are already in the image;
# hyperopt / seaborn / matplotlib are not, see ref_loader.py).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
rm -rf /tmp/refsrc "$HERE/_ref"
cp -r /root/reference /tmp/refsrc
sed -i "s/    setup_requires=\['sphinx>=2.1.2'\],//" /tmp/refsrc/setup.py
cd /tmp/refsrc
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" .
# data shipped with the reference that its installer does not package: the pretrained FB15k TransE
# checkpoint used by the drop-in test (Trainer.load_model, pykg2vec/utils/trainer.py:399-419)
mkdir -p "$HERE/_ref/examples/pretrained/TransE"
cp /root/reference/examples/pretrained/TransE/model.vec.pt /root/reference/examples/pretrained/TransE/config.npy \
   "$HERE/_ref/examples/pretrained/TransE/"
echo "installed: $(ls "$HERE/_ref")"
