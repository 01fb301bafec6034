"""Bench / test harness: workload definitions and run drivers.  Not imported by the product package."""
