"""Packs the reference's committed INPUT datasets the parity tests need
(/root/reference/dataset/{NonnegPCA,StableIdentification}/1; whitespace-separated
np.savetxt %.18e files, src/base/dataset_generator.py:39-50) into one JSON fixture,
tests/golden/datasets.json, so that tests run on the GPU box where /root/reference
does not exist.  Floats round-trip exactly through repr().

    python tests/golden/make_datasets.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/dataset"


def load(p):
    return np.loadtxt(p).tolist()


doc = {
    "NonnegPCA/1": {k: load(f"{REF}/NonnegPCA/1/{k}.csv") for k in ("dim", "Z", "initx_a", "initineqLagmult")},
    "StableIdentification/1": {},
}
si = doc["StableIdentification/1"]
for k in ["dim", "constset", "initineqLagmult"] + [f"noisyX_{i}" for i in range(1, 6)]:
    si[k] = load(f"{REF}/StableIdentification/1/{k}.csv")
for pt in "abcdefghijklmnopqrst":
    for c in "JRQ":
        si[f"init{c}_{pt}"] = load(f"{REF}/StableIdentification/1/init{c}_{pt}.csv")
with open(os.path.join(HERE, "datasets.json"), "w") as f:
    json.dump(doc, f)
print("wrote datasets.json", os.path.getsize(os.path.join(HERE, "datasets.json")))
