import numpy as np
import plaac_b200
from oracle import orc
from tests.test_gpu_fuzz import _random_case
kw, seqs, api = _random_case(17)
i = [k for k, s in enumerate(seqs) if len(s) == 8200][0]
codes, offs = plaac_b200.pack([seqs[i]])
P = orc.make_params(**kw)
ref = orc.score_batch(P, codes, offs)
print("ref", ref["hmm_all"][0].hex(), ref["hmm_vit"][0].hex())
for lm, warm in ((1024, 0), (1024, -1), (1024, 64), (1024, 255), (0, 0)):
    sc = plaac_b200.Scorer(plaac_b200.default_params(**kw))
    sc.set_long_path(lm, warm)
    got = sc.score(codes, offs)
    st = sc.stats()
    print(lm, warm, got["hmm_all"][0].hex(), got["hmm_vit"][0].hex(), "redone", st.long_redone_chunks, "long", st.long_proteins)
    sc.close()
