"""The oracle is test infrastructure: nothing under audiolazy_b200/ (nor the native
sources) may import, link, open or mention it, and the product has no CPU evaluator."""
import os
import re

from conftest import ROOT


def product_files():
  for base, _, names in os.walk(os.path.join(ROOT, "audiolazy_b200")):
    if "__pycache__" in base:
      continue
    for name in names:
      if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
        yield os.path.join(base, name)
  yield os.path.join(ROOT, "include", "alz_b200.h")


def test_product_never_touches_oracle():
  pattern = re.compile(r"(import\s+oracle|from\s+oracle|alz_oracle|libalz_oracle|orc_bank_apply|/oracle/)")
  offenders = [f for f in product_files() if pattern.search(open(f, encoding="utf-8").read())]
  assert not offenders, offenders


def test_product_has_no_reference_import():
  pattern = re.compile(r"(^|\s)(import\s+audiolazy\b|from\s+audiolazy\b|/root/reference)")
  offenders = [f for f in product_files() if pattern.search(open(f, encoding="utf-8").read())]
  assert not offenders, offenders


def test_product_has_no_scipy_filter_fallback():
  pattern = re.compile(r"(lfilter|sosfilt|scipy\.signal)")
  offenders = [f for f in product_files() if pattern.search(open(f, encoding="utf-8").read())]
  assert not offenders, offenders
