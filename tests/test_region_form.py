"""Gt/Lt on the device path: the region form of bounded requirements (include/ksched.h: ksched_key_regions) is checked
against the host algebra, which the reference's golden vectors pin (tests/test_golden_requirements.py)."""
from conftest import load_pkg


def test_region_form_commutes_with_the_host_algebra():
    k = load_pkg()
    assert k.lib().kh_region_selftest(1, 200000) == 0
    assert k.lib().kh_region_selftest(7, 200000) == 0
