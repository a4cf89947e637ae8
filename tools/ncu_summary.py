"""Prints the metrics we track from an .ncu-rep (raw page) as CSV rows: metric,unit,values..."""
import csv, subprocess, sys
KEEP = ['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_warps',
 'sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_evict_last_lookup_hit.sum',
 'lts__t_sectors_srcunit_tex_evict_last_lookup_miss.sum','lts__t_sectors_srcunit_tex_evict_first_lookup_miss.sum',
 'l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum','l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
 'sm__cycles_elapsed.max','smsp__average_warp_latency_per_inst_issued.ratio','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = KEEP + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    print('metric,unit,' + ','.join('launch%d' % i for i in range(len(rows) - 2)))
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            print('%s,%s,%s' % (k, units[i], ','.join(r[i].replace(',', ';')[:110] for r in rows[2:])))
if __name__ == '__main__':
    main(sys.argv[1])
