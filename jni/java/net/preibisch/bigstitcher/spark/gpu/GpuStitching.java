package net.preibisch.bigstitcher.spark.gpu;

import net.imglib2.Interval;
import net.imglib2.RandomAccessibleInterval;
import net.imglib2.img.array.ArrayImgs;
import net.imglib2.realtransform.Translation3D;
import net.imglib2.type.numeric.integer.UnsignedShortType;
import net.imglib2.util.Pair;
import net.imglib2.util.Util;
import net.imglib2.util.ValuePair;
import net.imglib2.view.Views;
import net.preibisch.stitcher.algorithm.PairwiseStitchingParameters;

/**
 * Drop-in for the numeric core of {@code TransformationTools.computeStitching}
 * (call site SparkPairwiseStitching.java:247-255): {@code PairwiseStitching.getShift( img1, img2, t1, t2, params, service )}
 * with the phase correlation, peak search and Pearson verification running in libbsgpu.
 *
 * Usage inside the per-pair task (replaces the body of TransformationTools.computeStitching after aggregation):
 * <pre>
 *   final Pair< Translation3D, Double > res = GpuStitching.getShift( ctx, input1, input2, t1, t2, params );
 *   if ( res == null ) return null;            // "No shift found" (SparkPairwiseStitching.java:274-279)
 * </pre>
 */
public final class GpuStitching
{
	private GpuStitching() {}

	/** overlap intervals as PairwiseStitching.getShift derives them; null when there is no equal-size overlap */
	public static long[][] localRasterOverlaps( final long[] dims1, final long[] dims2, final double[] t1, final double[] t2 )
	{
		final long[][] out = new long[ 3 ][ 3 ]; // min1, min2, size
		final double[] sub = new double[ 6 ];
		for ( int d = 0; d < 3; ++d )
		{
			final double lo = Math.max( t1[ d ], t2[ d ] ), hi = Math.min( t1[ d ] + dims1[ d ] - 1, t2[ d ] + dims2[ d ] - 1 );
			if ( hi < lo ) return null;
			final long a1 = (long)Math.ceil( lo - t1[ d ] - 1e-9 ), b1 = (long)Math.floor( hi - t1[ d ] + 1e-9 );
			final long a2 = (long)Math.ceil( lo - t2[ d ] - 1e-9 ), b2 = (long)Math.floor( hi - t2[ d ] + 1e-9 );
			if ( b1 - a1 + 1 <= 0 || b1 - a1 != b2 - a2 ) return null;
			out[ 0 ][ d ] = a1; out[ 1 ][ d ] = a2; out[ 2 ][ d ] = b1 - a1 + 1;
		}
		return out;
	}

	/**
	 * @return (shift of image 2 relative to image 1, cross correlation) or null
	 */
	public static Pair< Translation3D, Double > getShift(
			final long ctx,
			final RandomAccessibleInterval< UnsignedShortType > img1,
			final RandomAccessibleInterval< UnsignedShortType > img2,
			final double[] t1, final double[] t2,
			final PairwiseStitchingParameters params )
	{
		final long[][] ov = localRasterOverlaps( img1.dimensionsAsLongArray(), img2.dimensionsAsLongArray(), t1, t2 );
		if ( ov == null ) return null;
		final short[] c1 = crop( img1, ov[ 0 ], ov[ 2 ] ), c2 = crop( img2, ov[ 1 ], ov[ 2 ] );
		final int[] iparams = { params.peaksToCheck, params.doSubpixel ? 1 : 0, 10, 10, 10 };
		final double[] r = BsNative.pcmPair( ctx, c1, c2, ov[ 2 ], BsNative.U16, iparams, params.minOverlap );
		if ( r[ BsNative.R_FOUND ] == 0 ) return null;
		final int off = params.doSubpixel ? BsNative.R_SHIFT_SUB : BsNative.R_SHIFT_INT;
		final double[] shift = new double[ 3 ];
		for ( int d = 0; d < 3; ++d )
		{
			// real-vs-raster offset of the two crops (interval.min - localOverlap.min), as in PairwiseStitching.getShift
			final double lo = Math.max( t1[ d ], t2[ d ] );
			final double sub1 = ov[ 0 ][ d ] - ( lo - t1[ d ] ), sub2 = ov[ 1 ][ d ] - ( lo - t2[ d ] );
			shift[ d ] = r[ off + d ] - ( sub2 - sub1 );
		}
		return new ValuePair<>( new Translation3D( shift ), r[ BsNative.R_R ] );
	}

	/** dense x-fastest copy of an interval (the GPU path wants flat primitive arrays, like ArrayImg) */
	static short[] crop( final RandomAccessibleInterval< UnsignedShortType > img, final long[] min, final long[] size )
	{
		final long[] max = new long[ 3 ];
		for ( int d = 0; d < 3; ++d ) max[ d ] = img.min( d ) + min[ d ] + size[ d ] - 1;
		final long[] mn = new long[] { img.min( 0 ) + min[ 0 ], img.min( 1 ) + min[ 1 ], img.min( 2 ) + min[ 2 ] };
		final Interval iv = new net.imglib2.FinalInterval( mn, max );
		final short[] out = new short[ (int)( size[ 0 ] * size[ 1 ] * size[ 2 ] ) ];
		net.imglib2.util.ImgUtil.copy( Views.interval( img, iv ), ArrayImgs.unsignedShorts( out, size ) );
		return out;
	}
}
