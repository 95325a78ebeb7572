"""Smoke/demo entry with the reference's shape (ssr_eval/test.py:21-38)."""
from .eval import SSR_Eval_Helper, BasicTestee


class MyTestee(BasicTestee):
    def infer(self, x):
        return x


def test(test_data_root="./datasets/vctk_test"):
    helper = SSR_Eval_Helper(
        MyTestee(), test_name="unprocessed", input_sr=44100, output_sr=44100, evaluation_sr=48000,
        setting_fft={"cutoff_freq": [12000]}, save_processed_result=True, test_data_root=test_data_root)
    return helper.evaluate(limit_test_nums=10, limit_test_speaker=-1)
