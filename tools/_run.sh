timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for k in 2 3 4; do echo "ZM_NPMIN=$k $(DFF_ZM_NPMIN=$k timeout 200 python tools/quick_time.py 16 10 384 576 bf16 2>&1 | tail -1)"; done
echo "base $(timeout 200 python tools/quick_time.py 16 10 384 576 bf16 2>&1 | tail -1)"
