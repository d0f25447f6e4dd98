"""The driver's entry points: build() must pass on the CPU box (everything cross-compiles, the library loads and reports
the header's ABI version); smoke() needs the GPU."""
import pytest

import __graft_entry__ as entry


def test_build_passes():
    entry.build()


@pytest.mark.gpu
def test_smoke_passes():
    entry.smoke()
