package net.preibisch.bigstitcher.spark.gpu;

import java.util.List;

import net.imglib2.Interval;
import net.imglib2.algorithm.blocks.BlockSupplier;
import net.imglib2.realtransform.AffineTransform3D;
import net.imglib2.type.NativeType;
import net.imglib2.type.numeric.RealType;

/**
 * Drop-in for the {@code BlockSupplier<T>} that {@code BlkAffineFusion.initWithIntensityCoefficients(...)} returns
 * (call site SparkAffineFusion.java:602-615); {@code BlockAlgoUtils.arrayImg( supplier, interval )} (:620-627) calls
 * {@link #copy(Interval, Object)} with the primitive array it allocated, and the block is fused on the device.
 *
 * The views (resident volume handles from {@link BsNative#volumeUpload}, adjusted registrations times mipmap
 * transforms, blending parameters after FusionTools.adjustBlending) are prepared once per task; windowed views
 * (only the source cells a block touches, OverlappingBlocks / ViewUtil.findOverlappingBlocks) pass fullDims / windowMin.
 */
public class GpuBlockSupplier< T extends RealType< T > & NativeType< T > > implements BlockSupplier< T >
{
	private final long ctx;
	private final T type;
	private final long[] bbMin;
	private final int nViews;
	private final double[] models;   // n * 12, source pixel -> world, ascending ViewId order
	private final long[] handles;    // n * {volume, content}
	private final float[] blend;     // n * {border[3], range[3]}
	private final long[] windows;    // n * {fullDims[3], windowMin[3]} or null
	private final int[] iparams;     // {fusionType, interpolation, outDtype, blendLutN, outBigEndian}
	private final double[] dparams;  // {minIntensity, maxIntensity}

	public GpuBlockSupplier( final long ctx, final T type, final long[] bbMin, final List< AffineTransform3D > srcToWorld,
			final long[] handles, final float[] blend, final long[] windows, final int fusionType, final int outDtype,
			final double minIntensity, final double maxIntensity )
	{
		this.ctx = ctx;
		this.type = type;
		this.bbMin = bbMin.clone();
		this.nViews = srcToWorld.size();
		this.models = new double[ nViews * 12 ];
		for ( int i = 0; i < nViews; ++i )
			System.arraycopy( srcToWorld.get( i ).getRowPackedCopy(), 0, models, i * 12, 12 );
		this.handles = handles;
		this.blend = blend;
		this.windows = windows;
		this.iparams = new int[] { fusionType, 1, outDtype, 0, 0 };   // little-endian: the block goes into an ArrayImg
		this.dparams = new double[] { minIntensity, maxIntensity };
	}

	@Override
	public T getType() { return type; }

	@Override
	public int numDimensions() { return 3; }

	/** interval is zero-min inside the bounding box (blockMin/blockMax, SparkAffineFusion.java:620-624); dest: float[] / short[] / byte[] */
	@Override
	public void copy( final Interval interval, final Object dest )
	{
		final long[] min = new long[ 3 ], size = new long[ 3 ];
		for ( int d = 0; d < 3; ++d )
		{
			min[ d ] = interval.min( d ) + bbMin[ d ];
			size[ d ] = interval.dimension( d );
		}
		// a failure surfaces as RuntimeException: RetryTrackerSpark re-queues the block (RetryTrackerSpark.java:41-61)
		BsNative.fuseBlock( ctx, nViews, models, handles, blend, windows, min, size, iparams, dparams, dest );
	}

	@Override
	public BlockSupplier< T > threadSafe() { return this; }   // bs_ctx calls are serialised by the library

	@Override
	public BlockSupplier< T > independentCopy() { return this; }
}
