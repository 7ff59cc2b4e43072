/*
 * Table option:  ALTER TABLE ks.t WITH compaction = {'class': 'org.apache.cassandra.db.compaction.GpuSizeTieredCompactionStrategy'};
 * Resolved by reflection (CompactionParams.java:296-299). Bucketing, thresholds and task selection are SizeTieredCompactionStrategy's;
 * only the task that EXECUTES a chosen bucket changes (AbstractCompactionStrategy.getCompactionTask :206, and the three task factories
 * :183-204 that STCS implements by calling `new CompactionTask(cfs, txn, gcBefore)` directly).
 * Not compiled in the build image (no JDK); see B200C.java for the compile line.
 */
package org.apache.cassandra.db.compaction;

import java.util.ArrayList;
import java.util.Collection;
import java.util.Map;

import org.apache.cassandra.db.ColumnFamilyStore;
import org.apache.cassandra.db.lifecycle.LifecycleTransaction;
import org.apache.cassandra.io.sstable.format.SSTableReader;

public class GpuSizeTieredCompactionStrategy extends SizeTieredCompactionStrategy
{
    public GpuSizeTieredCompactionStrategy(ColumnFamilyStore cfs, Map<String, String> options)
    {
        super(cfs, options);
    }

    /** a stock task is kept whenever the GPU engine's envelope does not cover the table or the inputs (B200C_EUNSUPPORTED cases, decided up front) */
    private AbstractCompactionTask gpuOrStock(AbstractCompactionTask stock)
    {
        if (!(stock instanceof CompactionTask) || stock.getClass() != CompactionTask.class)
            return stock;
        LifecycleTransaction txn = stock.transaction;
        if (!GpuCompactionTask.supports(cfs, txn.originals()))
            return stock;
        return new GpuCompactionTask(cfs, txn, ((CompactionTask) stock).gcBefore);
    }

    @Override
    public AbstractCompactionTask getNextBackgroundTask(long gcBefore)
    {
        AbstractCompactionTask t = super.getNextBackgroundTask(gcBefore);
        return t == null ? null : gpuOrStock(t);
    }

    @Override
    public synchronized Collection<AbstractCompactionTask> getMaximalTask(long gcBefore, boolean splitOutput)
    {
        Collection<AbstractCompactionTask> tasks = super.getMaximalTask(gcBefore, splitOutput);
        if (tasks == null)
            return null;
        Collection<AbstractCompactionTask> out = new ArrayList<>(tasks.size());
        for (AbstractCompactionTask t : tasks)
            out.add(gpuOrStock(t));
        return out;
    }

    @Override
    public AbstractCompactionTask getUserDefinedTask(Collection<SSTableReader> sstables, long gcBefore)
    {
        AbstractCompactionTask t = super.getUserDefinedTask(sstables, gcBefore);
        return t == null ? null : gpuOrStock(t);
    }

    @Override
    public AbstractCompactionTask getCompactionTask(LifecycleTransaction txn, long gcBefore, long maxSSTableBytes)
    {
        return GpuCompactionTask.supports(cfs, txn.originals()) ? new GpuCompactionTask(cfs, txn, gcBefore)
                                                                 : super.getCompactionTask(txn, gcBefore, maxSSTableBytes);
    }
}
