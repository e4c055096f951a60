"""Known answers of the reference's utils/language_utils.py:tokens2description on seeded random token lists ->
tests/golden/detok_cases.json (run here, where /root/reference exists)."""
import importlib.util, json, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("ref_lu", "/root/reference/utils/language_utils.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
vocab = [f"w{i}" for i in range(40)]
vocab[3], vocab[4], vocab[5] = "Hello", "WORLD", "a"
random.seed(7)
cases = []
for _ in range(60):
    toks = [random.randrange(40) for _ in range(random.randrange(1, 14))]
    try:
        out = ref.tokens2description(toks, vocab, 0, 1)
    except IndexError:
        out = None
    cases.append({"tokens": toks, "caption": out})
json.dump({"vocab": vocab, "sos": 0, "eos": 1, "cases": cases}, open(os.path.join(ROOT, "tests", "golden", "detok_cases.json"), "w"))
print("wrote", len(cases))
