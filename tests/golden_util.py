import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = ["ivfpq_l2_d32_m8", "ivfpq_ip_d64_m32"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(HERE, "golden", name + ".npz"))
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        for k, v in self.meta.items():
            setattr(self, k, v)
        off = np.zeros(self.nlist + 1, np.int64)
        off[1:] = np.cumsum(z["list_lens"])
        self.lists = [(z["list_ids"][off[l]:off[l + 1]], z["list_codes"][off[l]:off[l + 1]]) for l in range(self.nlist)]
        self.filters = [(0, self.N - 1, False, z["filt_flags"]),
                        (int(z["filt2_lo"]), int(z["filt2_hi"]), True, z["filt2_flags"])]
        self.deleted = z["deleted"]
        self.window = (float(z["window"][0]), float(z["window"][1]))

    def __getitem__(self, k):
        return self.z[k]
