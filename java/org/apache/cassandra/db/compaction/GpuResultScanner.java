/*
 * ISSTableScanner view of a GPU compaction (SURVEY §8b, "Scanner" row): the native call bypasses the per-row pulls of the input scanners, but
 * callers that want an iterator over the MERGED result (scrub / verify tooling, tests that compare with a stock CompactionIterator) get one here:
 * the scanner of the output sstable the task just registered, with the accounting methods answering for the whole compaction
 * (S/io/sstable/ISSTableScanner.java:34-41). Progress DURING the call comes from B200C.poll / B200C.pollInputs (GpuCompactionTask.GpuCompactionInfo).
 * Not compiled in the build image (no JDK); see B200C.java for the compile line.
 */
package org.apache.cassandra.db.compaction;

import java.util.Collection;
import java.util.Set;

import org.apache.cassandra.db.rows.UnfilteredRowIterator;
import org.apache.cassandra.io.sstable.ISSTableScanner;
import org.apache.cassandra.io.sstable.format.SSTableReader;
import org.apache.cassandra.schema.TableMetadata;

public final class GpuResultScanner implements ISSTableScanner
{
    private final ISSTableScanner merged;           // scanner over the output sstable
    private final Collection<SSTableReader> inputs;
    private final long inputBytes, inputCompressedBytes;

    public GpuResultScanner(SSTableReader output, Collection<SSTableReader> inputs)
    {
        this.merged = output.getScanner();
        this.inputs = inputs;
        long u = 0, c = 0;
        for (SSTableReader r : inputs) { u += r.uncompressedLength(); c += r.onDiskLength(); }
        this.inputBytes = u; this.inputCompressedBytes = c;
    }

    @Override public long getLengthInBytes() { return inputBytes; }
    @Override public long getCompressedLengthInBytes() { return inputCompressedBytes; }
    @Override public long getCurrentPosition() { return merged.getCurrentPosition(); }
    @Override public long getBytesScanned() { return inputBytes; }          // the compaction has consumed every input byte by the time this exists
    @Override public Set<SSTableReader> getBackingSSTables() { return Set.copyOf(inputs); }
    @Override public TableMetadata metadata() { return merged.metadata(); }
    @Override public boolean hasNext() { return merged.hasNext(); }
    @Override public UnfilteredRowIterator next() { return merged.next(); }
    @Override public void close() { merged.close(); }
}
