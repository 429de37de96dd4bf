"""Generate tests/golden/ref_blocks/*.pb by running the REFERENCE's own block writer
(`barrier_array_merge`, /root/reference/utils/util.py:88-143, called the way
drivers/gen_passage_embeddings.py:146-169 `StreamInferenceDoc` calls it), imported in this container
with stub modules for what the image lacks (pytrec_eval, the transformers-2.3 model code) and with
`torch.distributed.barrier` neutralised (single process).  The files pin the on-disk block format that
convdr_b200/blocks.py writes and reads.

Run from the repo root (needs /root/reference; the GPU box does not have it — only the .pb files travel):
    python tests/golden/make_golden_blocks.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from convdr_b200 import blocks, synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_blocks")
RANK, N_ROWS, WORLD, SEED = 3, 5, 8, 17


def load_reference_util():
    stubs = {}
    for name in ["pytrec_eval", "utils", "utils.dpr_utils", "model", "model.models"]:
        stubs[name] = types.ModuleType(name)
    stubs["utils.dpr_utils"].get_model_obj = None
    stubs["utils.dpr_utils"].load_states_from_checkpoint = None
    stubs["model.models"].MSMarcoConfigDict = {}
    stubs["model.models"].ALL_MODELS = ()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_utils_util", os.path.join(REF, "utils", "util.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    mod = load_reference_util()
    mod.dist.barrier = lambda *a, **k: None            # single process: nothing to wait for
    os.makedirs(OUT, exist_ok=True)
    emb = synth.block(0, N_ROWS, seed=SEED)
    embid = blocks.strided_offsets(N_ROWS * WORLD, RANK, WORLD)
    args = types.SimpleNamespace(local_rank=0, rank=RANK, output_dir=OUT, world_size=WORLD)
    prefix = "passage_"                                 # gen_passage_embeddings.py: StreamInferenceDoc(..., "passage_", ...)
    mod.barrier_array_merge(args, emb, prefix=prefix + "_emb_p_", load_cache=False, only_load_in_master=True, merge=False)
    mod.barrier_array_merge(args, embid, prefix=prefix + "_embid_p_", load_cache=False, only_load_in_master=True, merge=False)
    with open(os.path.join(OUT, "params.json"), "w") as f:
        json.dump({"rank": RANK, "n_rows": N_ROWS, "world": WORLD, "seed": SEED, "numpy": np.__version__,
                   "files": sorted(x for x in os.listdir(OUT) if x.endswith(".pb"))}, f, indent=1)
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
