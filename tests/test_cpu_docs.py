"""Documentation that is cheap to keep honest: every environment switch the library reads is listed in DESIGN.md's
appendix, and every profile file DESIGN.md cites exists."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts), encoding="utf-8") as f:
        return f.read()


def test_every_library_switch_is_documented():
    design = _read("DESIGN.md")
    csrc = os.path.join(ROOT, "minsdtf_b200", "csrc")
    switches = set()
    for name in os.listdir(csrc):
        switches |= set(re.findall(r'(?:getenv|env_int)\("(SDTF_[A-Z0-9_]+)"', _read("minsdtf_b200", "csrc", name)))
    assert len(switches) > 20
    missing = sorted(s for s in switches if s not in design)
    assert not missing, f"switches read by libsdtf.so but absent from DESIGN.md: {missing}"


def test_cited_profiles_exist():
    design = _read("DESIGN.md")
    cited = set(re.findall(r"profiles/([A-Za-z0-9_.\-]+\.[a-z]+)", design))
    cited |= set(re.findall(r"`(r0[12]_[A-Za-z0-9_.\-]+\.(?:log|md|json|jsonl|txt|csv))`", design))  # bare `r02_x_name.log`
    assert len(cited) > 20
    missing = sorted(f for f in cited if not os.path.exists(os.path.join(ROOT, "profiles", f)))
    assert not missing, f"DESIGN.md cites profile files that are not committed: {missing}"
