"""Driver for the round-2 ncu captures of the kernels the bench-step capture does not reach (development tool):
the scoring head (rotary GEMM epilogue, LSE GEMM epilogue + finish, device tokenise/pad) and the generation loop
(sampler step, counters).  Two warm passes, then the pass the -s/-c window of the ncu command picks.
    ncu --set full --clock-control none -k regex:<pattern> -s <n> -c <m> -o gpurun_out/<name> python tools/ncu_targets_r02.py score|gen"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import evo_b200  # noqa: E402
from evo_b200.models import load_checkpoint  # noqa: E402

what = sys.argv[1]
dev = "cuda:0"
tok = evo_b200.CharLevelTokenizer(512)
rng = np.random.default_rng(0)
model = load_checkpoint("evo-1-8k-base", device=dev, random_init=True, seed=0)
if what == "score":
    seqs = ["".join(rng.choice(list("ACGT"), size=8192)) for _ in range(8)]
    for _ in range(4):            # per pass the library launches: tokenize_pad, 129 x gemm_tcgen05_kernel (the last one = LSE epilogue), score_finish
        evo_b200.score_sequences(seqs, model, tok, device=dev)
elif what == "gen":
    seqs = ["".join(rng.choice(list("ACGT"), size=1024)) for _ in range(16)]
    evo_b200.generate(seqs, model, tok, n_tokens=24, top_k=4, top_p=0.9, temperature=0.9, cached_generation=True, verbose=0, device=dev, force_prompt_threshold=1024)
torch.cuda.synchronize()
print("done", what)
