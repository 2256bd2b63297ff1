"""Extracts the reference's numeric comparison truth table (tidb_query_expr/src/impl_compare.rs
`generate_numeric_compare_cases`, used by test_compare_real / test_compare_duration / test_compare_decimal) into
tests/golden/compare_cases.json.  Run in the build container, where /root/reference exists; the JSON travels."""
import json
import os
import re

SRC = "/root/reference/components/tidb_query_expr/src/impl_compare.rs"
text = open(SRC).read()
start = text.index("fn generate_numeric_compare_cases()")
body = text[start:text.index("fn test_compare_real()", start)]
first_line = text[:start].count("\n") + 1
arg = r"(None|Real::new\((-?[0-9.]+)\)\.ok\(\))"
pat = re.compile(r"\(\s*" + arg + r",\s*" + arg + r",\s*TestCaseCmpOp::(\w+),\s*(None|Some\((\d)\)),?\s*\)", re.S)
cases = []
for m in pat.finditer(body):
    a = None if m.group(1) == "None" else float(m.group(2))
    b = None if m.group(3) == "None" else float(m.group(4))
    cases.append({"a": a, "b": b, "op": m.group(5), "expect": None if m.group(6) == "None" else int(m.group(7))})
assert len(cases) == 63, len(cases)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compare_cases.json")
json.dump({"source": f"impl_compare.rs:{first_line}-{first_line + body.count(chr(10))} generate_numeric_compare_cases (63 rows)", "cases": cases}, open(out, "w"), indent=0)
print(len(cases), "cases ->", out)
