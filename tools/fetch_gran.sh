M=dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/fetch_gran_default.csv tools/fetch_gran > gpurun_out/fetch_gran_default.jsonl 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/fetch_gran_lim32.csv tools/fetch_gran 32 > gpurun_out/fetch_gran_lim32.jsonl 2>&1
tools/fetch_gran > gpurun_out/fetch_gran_plain.jsonl 2>&1
tools/fetch_gran 32 > gpurun_out/fetch_gran_plain32.jsonl 2>&1
