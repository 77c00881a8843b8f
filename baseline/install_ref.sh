#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with gpurun).
# The reference's setup.py has a packaging bug -- find_packages(where='contrastors') with package_dir={'': 'src'} finds
# nothing, so `pip install /root/reference` yields an empty wheel (the README installs with `pip install -e .`, which only
# adds src/ to the path).  The install therefore runs from a /tmp copy whose setup.py says where='src'; no module under
# src/contrastors is touched.  --no-deps: deepspeed / megablocks / wandb-era pins are not in the offline wheelhouse; the
# modules bench.py imports (loss, distributed, rand_state, models.huggingface.*) need torch / transformers / einops only.
set -e
cd "$(dirname "$0")/.."
[ -d /root/reference ] || { echo "no /root/reference here (GPU box): keeping the installed baseline/_ref"; exit 0; }
rm -rf /tmp/refcopy baseline/_ref
cp -r /root/reference /tmp/refcopy
sed -i "s/find_packages(where='contrastors')/find_packages(where='src')/" /tmp/refcopy/setup.py
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/refcopy 2>&1 | tail -2
# byte-for-byte check of every installed module against the read-only tree
(cd baseline/_ref/contrastors && find . -name '*.py' | sort | while read f; do cmp -s "$f" "/root/reference/src/contrastors/$f" || { echo "MODIFIED: $f"; exit 1; }; done)
echo "baseline/_ref: $(find baseline/_ref/contrastors -name '*.py' | wc -l) modules, identical to /root/reference/src/contrastors"
