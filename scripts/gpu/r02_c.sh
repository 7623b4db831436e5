python -m pytest tests/test_host_cli.py tests/test_ingest.py -m gpu -x -q 2>&1 | tail -5
bash scripts/gpu/r02_cli_time.sh
PLAAC_CLI_TIMING=1 plaac_b200/bin/plaac -i /tmp/huge.fa > /tmp/out.tsv
