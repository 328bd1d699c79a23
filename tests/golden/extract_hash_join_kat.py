#!/usr/bin/env python
"""Transcribe the reference's HashJoinExec unit-test fixtures into tests/golden/hash_join_kat.json.

Source: /root/reference/datafusion/physical-plan/src/joins/hash_join/exec.rs (test module, :2779-8136).
Each extracted case records the reference line of the test, the literal input tables, the join
keys / type / null-equality and the insta snapshot of the expected output (exact order when the
reference asserts with `batches_to_string`, order-insensitive when it uses `batches_to_sort_string`).
The reference runs every such test at batch_size in {8192,10,5,2,1} x perfect-hash {on,off}
(exec.rs:2929-2962); tests/test_golden_join.py mirrors that matrix.

Run (only where /root/reference exists):  python tests/golden/extract_hash_join_kat.py
"""
import json
import os
import re
import sys

SRC = "/root/reference/datafusion/physical-plan/src/joins/hash_join/exec.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hash_join_kat.json")


def parse_vec(s):
    """'vec![1, 2, 3]' or 'vec![Some(1), None]' -> list with None for NULL"""
    inner = s[s.index("[") + 1:s.rindex("]")]
    out = []
    for tok in re.split(r",\s*", inner.strip().rstrip(",")):
        tok = tok.strip()
        if not tok:
            continue
        if tok == "None":
            out.append(None)
        else:
            m = re.match(r"Some\((-?\d+)\)", tok)
            out.append(int(m.group(1)) if m else int(re.sub(r"_?i(32|64)$", "", tok)))
    return out


def parse_table_call(expr):
    """build_table(("a1", &vec![..]), ...) or build_table_two_cols(...) -> [(name, values)]"""
    cols = re.findall(r'\(\s*"(\w+)"\s*,\s*&(vec!\[[^\]]*\])\s*,?\s*\)', expr, re.S)
    return [(n, parse_vec(v)) for n, v in cols]


def balanced(text, start):
    """text[start] == '(' -> index after the matching ')'"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def main():
    text = open(SRC).read()
    lines = text.split("\n")
    # helper fixtures: fn build_xxx() -> Arc<dyn ExecutionPlan> { build_table(...) }
    helpers = {}
    for m in re.finditer(r"fn (build_\w+)\(\)\s*->\s*Arc<dyn ExecutionPlan>\s*\{", text):
        body_start = m.end()
        call = text.find("build_table(", body_start)
        if call < 0 or call - body_start > 400:
            continue
        end = balanced(text, text.index("(", call))
        helpers[m.group(1)] = parse_table_call(text[call:end])
    cases = []
    fn_iter = list(re.finditer(r"async fn (\w+)\(", text))
    for k, m in enumerate(fn_iter):
        name = m.group(1)
        start = m.start()
        end = fn_iter[k + 1].start() if k + 1 < len(fn_iter) else len(text)
        body = text[start:end]
        line = text.count("\n", 0, start) + 1
        if line < 2779:
            continue
        if "for join_type in" in body:  # multi-join-type loops are transcribed by hand into misc_kat.json
            continue
        # join_*_with_filter tests share prepare_join_filter() (exec.rs:5556-5583): left.c@2 > right.c@2
        shared_filter = "prepare_join_filter()" in body and "JoinFilter::new" not in body
        if ("JoinFilter" in body or "filter" in name) and not shared_filter:
            continue
        if "struct" in name or "dict" in name or "null_aware" in name:
            continue
        tables = {}
        repeat = {"left": 1, "right": 1}
        ok = True
        for side in ("left", "right"):
            mm = re.search(r"let %s = (\w+)\(" % side, body)
            if not mm:
                ok = False
                break
            fn = mm.group(1)
            if fn in ("build_table", "build_table_two_cols", "build_table_two_batches"):
                e = balanced(body, body.index("(", mm.start()))
                tables[side] = parse_table_call(body[mm.start():e])
                repeat[side] = 2 if fn == "build_table_two_batches" else 1   # exec.rs:3853-3861: the same batch twice
            elif fn in helpers:
                tables[side] = helpers[fn]
            else:
                ok = False
                break
            if not tables.get(side):
                ok = False
        if not ok:
            continue
        on = re.findall(r'Column::new_with_schema\("(\w+)",\s*&(left|right)\.schema\(\)\)', body)
        if not on or len(on) % 2:
            continue
        pairs = []
        for i in range(0, len(on), 2):
            if on[i][1] != "left" or on[i + 1][1] != "right":
                ok = False
            pairs.append([on[i][0], on[i + 1][0]])
        if not ok:
            continue
        jt = re.search(r"JoinType::(\w+)", body)
        ne = re.search(r"NullEquality::(\w+)", body)
        snap = re.search(r'assert_snapshot!\(\s*(batches_to_string|batches_to_sort_string)\(&batches\),\s*@r"(.*?)"\s*\)', body, re.S)
        sorted_cmp = None
        if snap:
            rows = [l.strip() for l in snap.group(2).split("\n") if l.strip().startswith("|")]
            sorted_cmp = snap.group(1) == "batches_to_sort_string"
        else:  # let expected = ["+--+", "| a |", ...]; assert_batches[_sorted]_eq!(expected, &batches)
            arr = re.search(r'let expected = \[(.*?)\];\s*assert_batches(_sorted)?_eq!\(expected, &batches\)', body, re.S)
            if arr:
                rows = [q.strip() for q in re.findall(r'"([^"]*)"', arr.group(1)) if q.strip().startswith("|")]
                sorted_cmp = arr.group(2) is not None
        if not jt or sorted_cmp is None or not rows:
            continue
        header = [c.strip() for c in rows[0].strip("|").split("|")]
        exp = []
        for r in rows[1:]:
            cells = [c.strip() for c in r.strip("|").split("|")]
            exp.append([None if c == "" else (True if c == "true" else False if c == "false" else int(c)) for c in cells])
        partitioned = "partitioned_join_collect" in body or "PartitionMode::Partitioned" in body
        cases.append(dict(name=name, ref="datafusion/physical-plan/src/joins/hash_join/exec.rs:%d" % line, left=tables["left"], right=tables["right"],
                          on=pairs, join_type=jt.group(1), null_equality=ne.group(1) if ne else "NullEqualsNothing",
                          sorted=sorted_cmp, partitioned=partitioned, header=header, expected=exp,
                          left_repeat=repeat["left"], right_repeat=repeat["right"],
                          filter=({"col_side": [0, 1], "col_index": [2, 2], "op": "gt", "ref": "exec.rs:5556-5583 prepare_join_filter"} if shared_filter else None)))
    json.dump(dict(source=SRC, note="transcribed by tests/golden/extract_hash_join_kat.py; do not edit by hand", cases=cases), open(OUT, "w"), indent=1)
    print("wrote %d cases to %s" % (len(cases), OUT))
    for c in cases:
        print("  %-55s %-10s rows=%d sorted=%s part=%s" % (c["name"], c["join_type"], len(c["expected"]), c["sorted"], c["partitioned"]))


if __name__ == "__main__":
    if not os.path.exists(SRC):
        sys.exit("reference tree not present: golden file is committed, nothing to do")
    main()
