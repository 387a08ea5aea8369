"""tests/test_api_gpu.py -- the unchanged histbook API on the device backend -- once more on the EMULATED engine
(tests/emul_engine.py): every test of that file that fills from host columns runs here on CPU, so the glue of
histbook_b200/__init__.py (Hist.fill / Book.fill installed into the reference's classes, `_content` over a Store, copy-on-fill,
views, + and *, group merges, pickle / JSON, project) is exercised where there is no GPU.  The
generated fill kernels really execute (on CPU threads); the engine's utility kernels (axpy, slice, table, project) are numpy
restatements in the stand-in, i.e. what is checked for those is the Python above them, not the CUDA below."""
import shutil

import pytest

histbook_b200 = pytest.importorskip("histbook_b200")
if not histbook_b200._ref.available():
    pytest.skip("the histbook package (baseline/_ref) is not available", allow_module_level=True)

import emul_engine                    # noqa: E402
import test_api_gpu as api            # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")

NEEDS_A_DEVICE = {"test_device_columns_pandas_and_read_side", "test_columns_produced_on_a_side_stream", "test_column_on_another_gpu_is_refused",
                  "test_cuda_array_interface_stream_key_is_honoured", "test_libm_class_functions_at_ten_million_events"}
TOO_MANY_EVENTS = {"test_table_select_and_fraction_on_the_device", "test_randomised_book_against_the_reference_path"}    # 1.5 - 2.5 minutes each on CPU threads (they pass)
NAMES = sorted(n for n in dir(api) if n.startswith("test_") and n not in NEEDS_A_DEVICE and n not in TOO_MANY_EVENTS)


@pytest.mark.parametrize("name", NAMES)
def test_histbook_api_on_the_emulated_engine(name):
    with emul_engine.installed():
        getattr(api, name)()
