"""TEST INFRASTRUCTURE — digest of the reference's character tables (needs /root/reference; run in the build container):

    python oracle/make_table_digest.py        -> tests/golden/character_table_digest.json

For every (group, signature) of the merged table (packaged `spaces/precomputed_characters.json` ∪ `plots/precomputed_characters.json`,
460 irreps) the file stores sha256 of the canonical form `sorted((exponents..., coefficient))`.  `tests/test_characters.py`
recomputes the same digest from `liestationarykernels_b200.characters` wherever it runs (the GPU box has no reference tree)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def canonical_digest(coeffs, monoms) -> str:
    rows = sorted(tuple(int(e) for e in m) + (int(c),) for c, m in zip(coeffs, monoms))
    return hashlib.sha256(json.dumps(rows).encode()).hexdigest()


def main():
    table = ref_shim.merged_reference_table()
    out = {}
    for group in sorted(table):
        out[group] = {sig: canonical_digest(*entry) for sig, entry in sorted(table[group].items())}
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "character_table_digest.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote", path, sum(len(v) for v in out.values()), "irreps")


if __name__ == "__main__":
    main()
