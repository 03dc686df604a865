#!/bin/bash
# Builds experiment variants of libdeeppde_b200.so (helper-group layouts of the tensor kernels) into deeppde_actorcritic_b200/variants/<name>.so;
# run one with DPB_LIB_PATH=deeppde_actorcritic_b200/variants/<name>.so.  usage: tools/variant_builds.sh name:VAR=val,-DMACRO=val ...
set -e
cd "$(dirname "$0")/.."
mkdir -p deeppde_actorcritic_b200/variants
for spec in "$@"; do
    name=${spec%%:*}; vars=${spec#*:}
    env DPB_EXTRA_FLAGS="$(echo "$vars" | tr ',' '\n' | grep '^-D' | tr '\n' ' ')" $(echo "$vars" | tr ',' '\n' | grep -v '^-D' | tr '\n' ' ') python __graft_entry__.py build > /dev/null
    cp deeppde_actorcritic_b200/libdeeppde_b200.so deeppde_actorcritic_b200/variants/$name.so
    echo "built $name ($vars)"
done
python __graft_entry__.py build > /dev/null      # back to the default build
