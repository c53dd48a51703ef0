import sys, torch, pytest
x = torch.full((1 << 28,), float("nan"), device="cuda"); del x
sys.exit(pytest.main(["tests/test_gpu_dp.py", "-x", "-q", "-m", "gpu", "-k", "emulated"]))
