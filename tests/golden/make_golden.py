"""Generates tests/golden/*.json from the reference tree (run in the dev container,
where /root/reference exists; the GPU box only sees the committed JSON).

Carried over (SURVEY.md 8(c) "Golden fixtures"):
  * the 10 rows of eg/fake_logit.c:12-22 with the canned fit of :38-40 and the true
    MLE recorded in SURVEY F9;
  * NIST StRD Pontius and Wampler1 data (tests/pontius.dat, tests/wampler1.dat) with
    the certified values asserted in tests/nist_tests.c:20-47;
  * NIST NumAcc4 (tests/numacc4.dat) with tests/nist_tests.c:49-55's certified mean/variance;
  * the small exact apop_map_sum / apop_dot cases of tests/test_apop.c:1037-1044, :527-550;
  * the real-data logit example: tests/amash_vote_analysis.csv (422 House votes on the Amash amendment,
    public data compiled by govtrack.us) as eg/logit.c:14-18 uses it -- vote ~ ideology + log(contribs+10) + party.
Only numbers are copied (public NIST data and test constants), no source text.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def table(path, skip=1):
    rows = []
    for ln in open(path).read().splitlines()[skip:]:
        f = ln.split()
        if f:
            rows.append([float(x) for x in f])
    return rows


def fake_logit():
    src = open(os.path.join(REF, "eg", "fake_logit.c")).read()
    body = src[src.index("outcome,A, B"):src.index('");', src.index("outcome,A, B"))]
    rows = []
    for ln in body.splitlines()[1:]:
        nums = re.findall(r"-?\d*\.?\d+", ln)
        if len(nums) == 3:
            rows.append([float(x) for x in nums])
    canned = [float(x) for x in re.findall(r"- (-?\d+\.\d+)\) < 1e-6", src)]
    assert len(rows) == 10 and len(canned) == 3
    return {"rows_outcome_A_B": rows, "reference_canned_fit_1_A_B": canned, "canned_tolerance": 1e-6,
            "true_mle_1_A_B": [-1.15502464, 4.03989415, 1.4946963],
            "source": "eg/fake_logit.c:12-22,38-40; true MLE from SURVEY.md F9"}


def nist():
    return {
        "pontius": {"y_x": table(os.path.join(REF, "tests", "pontius.dat")),
                    "certified_beta": [0.673565789473684E-03, 0.732059160401003E-06, -0.316081871345029E-14],
                    "certified_se": [0.107938612033077E-03, 0.157817399981659E-09, 0.486652849992036E-16],
                    "tolerances": [1e-9, 1e-15, 1e-15], "source": "tests/pontius.dat, tests/nist_tests.c:20-35"},
        "wampler1": {"y_x": table(os.path.join(REF, "tests", "wampler1.dat")),
                     "certified_beta": [1, 1, 1, 1, 1, 1], "tolerance": 1e-3,
                     "source": "tests/wampler1.dat, tests/nist_tests.c:37-47"},
        "numacc4": {"x": [r[0] for r in table(os.path.join(REF, "tests", "numacc4.dat"), skip=0)],
                    "certified_mean": 10000000.2, "certified_population_variance": 0.01,
                    "source": "tests/numacc4.dat, tests/nist_tests.c:49-55"},
    }


def amash():
    rows = []
    for ln in open(os.path.join(REF, "tests", "amash_vote_analysis.csv")).read().splitlines():
        f = ln.strip().split(",")
        if len(f) == 5 and f[0].isdigit():
            rows.append({"id": int(f[0]), "party": f[1], "ideology": float(f[2]), "vote": f[3], "contribs": float(f[4])})
    assert len(rows) == 422
    return {"rows": rows, "source": "tests/amash_vote_analysis.csv; model of eg/logit.c:14-23: select 0, ideology, "
                                    "log(contribs+10), vote, party; apop_data_to_factors (vote -> column 0); "
                                    "apop_data_to_dummies(.col=1, .type='t', .append='y') (party); apop_estimate(d, apop_logit)"}


if __name__ == "__main__":
    json.dump(fake_logit(), open(os.path.join(HERE, "fake_logit.json"), "w"), indent=1)
    json.dump(nist(), open(os.path.join(HERE, "nist_strd.json"), "w"))
    json.dump(amash(), open(os.path.join(HERE, "amash.json"), "w"))
    print("wrote fake_logit.json, nist_strd.json, amash.json")
