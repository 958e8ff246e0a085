"""Markdown tables from the output of configs_bench.py:
    python profiles/scripts/configs_table.py gpurun_out/r02ae_configs.jsonl > profiles/r02_configs.md"""
import json
import sys

rows = [json.loads(x) for x in open(sys.argv[1]) if x.strip().startswith("{")]
srch = [r for r in rows if r["config"] != "C5 all-pairs"]
allp = [r for r in rows if r["config"] == "C5 all-pairs"]
print("# r02: the other BASELINE.json configurations (one B200)\n")
print("Script: `profiles/scripts/configs_bench.py` (parity of every configuration is in `tests/`; this records speed).")
print("Device time = CUDA-event time of hash + index build + search stages (the script's hit buffer is pageable numpy")
print("memory, so its copy-out is excluded; `qhash` includes the host-side embedding and upload of the 10 k query strings).")
print("`path`: dense u16 bucket ranks, packed 64-bit keys, or (KW >= 2) the sort on a 64-bit hash of the key.\n")
print("## LSH search: configs[0] (C1) and the K x L x W sweep of configs[1] at 10 M fragments, 10 k queries, len 10, R = 30\n")
print("| config | N | Q | K | L | W | path | KW | sort passes | candidates | hits | hash ms | build ms (sort / group / stores) | search ms (qhash / filter / exact / hit sort) | device ms | DB fragments/s |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in srch:
    path = "ranks" if r["rank_path"] else ("key hash" if r["key_words"] >= 2 and r["sort_passes"] <= 8 * r["L"] else "keys")
    print(f"| {r['config']} | {r['n_db']:,} | {r['n_query']:,} | {r['K']} | {r['L']} | {r['W']:g} | {path} | {r['key_words']} | "
          f"{r['sort_passes']} | {r['candidates']:.3g} | {r['hits']:,} | {r['hash_ms']} | {r['build_ms']} ({r.get('sort_ms', '')} / "
          f"{r.get('group_ms', '')} / {r.get('permute_ms', '')}) | {r['search_ms']} ({r.get('qhash_ms', '')} / {r['filter_ms']} / "
          f"{r['exact_ms']} / {r['hitsort_ms']}) | {r['device_ms']} | {r['db_fragments_per_s']:.3g} |")
print("\n## configs[4] (C5): all pairs of 1 M fragments (4.99995e11 pairs), hits only\n")
print("| len | metric | R | survivors | hits | device ms | pairs/s | filter ms | exact ms |")
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|")
for r in allp:
    print(f"| {r['len']} | {r['metric']} | {r['R']:g} | {r['survivors']:,} | {r['hits']:,} | {r['device_ms']} | {r['pairs_per_s']:.3g} | "
          f"{r['filter_ms']} | {r['exact_ms']} |")
