#!/usr/bin/env python3
"""Extract golden vectors for the hot path from the reference's own test sources.

Run in the build container (where /root/reference exists); the JSON it writes is committed so the
tests never need /root/reference at run time. Only literal expected values are extracted
(integer tables and closed-form constants) -- no reference code is copied.

  vertex_inds.json : every `hexed::vertex_inds(n_dim, {{i_dim0, i_dim1}, {sign0, sign1}})` case with the
                     REQUIREd entries of test/test_connection.cpp:14-131
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))

text = open(os.path.join(ref, "test", "test_connection.cpp")).read()
block = text[text.index('TEST_CASE("vertex_inds")'):text.index('TEST_CASE("Element_face_connection<Element>")')]
cases = []
pieces = re.split(r"hexed::vertex_inds\(", block)[1:]
for piece in pieces:
    m = re.match(r"(\d), \{\{(\d), (\d)\}, \{(\d), (\d)\}\}\);", piece)
    nd, d0, d1, s0, s1 = map(int, m.groups())
    n_vert = 2**(nd - 1)
    inds = [[None]*n_vert, [None]*n_vert]
    for side, i, val in re.findall(r"REQUIRE\(inds\[(\d)\]\[(\d)\] == (\d)\)", piece):
        inds[int(side)][int(i)] = int(val)
    assert all(v is not None for row in inds for v in row)
    cases.append({"n_dim": nd, "i_dim": [d0, d1], "face_sign": [s0, s1], "inds": inds})
json.dump({"source": "test/test_connection.cpp:14-131", "cases": cases}, open(os.path.join(here, "vertex_inds.json"), "w"), indent=1)
print(len(cases), "vertex_inds cases")
