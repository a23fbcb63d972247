// Pins the oracle: runs the REAL upstream arithmetic (BigStitcher 2.5.0 / multiview-reconstruction 8.0.0, the versions
// pom.xml:106-107 of the reference pins) on the seeded inputs of tests/golden/make_jvm_inputs.py and writes
// tests/golden/jvm/{pcm_*.json, pcm_*_pcm.raw, fusion_*_<TYPE>.raw, dog_*.json}.  tests/test_jvm_golden.py consumes them.
//
//   cd tests/golden/java && mvn -q compile exec:java -Dexec.args="../jvm_inputs ../jvm"
//
// NOT compiled in the build image (no JDK / Maven, no network): written against the public upstream APIs named at the
// reference's call sites (SparkPairwiseStitching.java:247-255, SparkAffineFusion.java:602-627,
// SparkInterestPointDetection.java:550-566).
import java.io.*;
import java.nio.*;
import java.nio.file.*;
import java.util.*;
import java.util.concurrent.*;

import com.google.gson.*;

import mpicbg.spim.data.generic.sequence.*;
import mpicbg.spim.data.sequence.*;
import net.imglib2.*;
import net.imglib2.algorithm.blocks.*;
import net.imglib2.algorithm.phasecorrelation.*;
import net.imglib2.img.array.*;
import net.imglib2.realtransform.*;
import net.imglib2.type.numeric.integer.UnsignedShortType;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.util.*;
import net.preibisch.mvrecon.fiji.plugin.fusion.FusionGUI.FusionType;
import net.preibisch.mvrecon.process.fusion.blk.BlkAffineFusion;
import net.preibisch.mvrecon.process.interestpointdetection.methods.dog.DoGImgLib2;
import net.preibisch.mvrecon.fiji.spimdata.interestpoints.InterestPoint;
import net.preibisch.stitcher.algorithm.PairwiseStitching;
import net.preibisch.stitcher.algorithm.PairwiseStitchingParameters;

public class GoldenDump
{
	static short[] readU16( final Path p, final int n ) throws IOException
	{
		final ShortBuffer sb = ByteBuffer.wrap( Files.readAllBytes( p ) ).order( ByteOrder.LITTLE_ENDIAN ).asShortBuffer();
		final short[] a = new short[ n ];
		sb.get( a );
		return a;
	}

	static void writeF32( final Path p, final float[] a ) throws IOException
	{
		final ByteBuffer bb = ByteBuffer.allocate( a.length * 4 ).order( ByteOrder.LITTLE_ENDIAN );
		bb.asFloatBuffer().put( a );
		Files.write( p, bb.array() );
	}

	public static void main( final String[] args ) throws Exception
	{
		final Path in = Paths.get( args[ 0 ] ), out = Paths.get( args[ 1 ] );
		Files.createDirectories( out );
		final JsonObject man = JsonParser.parseReader( Files.newBufferedReader( in.resolve( "manifest.json" ) ) ).getAsJsonObject();
		final ExecutorService service = Executors.newFixedThreadPool( 1 );
		final Gson gson = new GsonBuilder().setPrettyPrinting().create();

		// ---- hot path 1: PairwiseStitching.getShift + the padded PCM
		for ( final JsonElement e : man.getAsJsonArray( "pcm" ) )
		{
			final JsonObject c = e.getAsJsonObject();
			final long[] dims = gson.fromJson( c.get( "dims" ), long[].class );
			final int n = (int)( dims[ 0 ] * dims[ 1 ] * dims[ 2 ] );
			final RandomAccessibleInterval< UnsignedShortType > a = ArrayImgs.unsignedShorts( readU16( in.resolve( c.get( "a" ).getAsString() ), n ), dims );
			final RandomAccessibleInterval< UnsignedShortType > b = ArrayImgs.unsignedShorts( readU16( in.resolve( c.get( "b" ).getAsString() ), n ), dims );
			final PairwiseStitchingParameters params = new PairwiseStitchingParameters();
			params.peaksToCheck = c.get( "peaksToCheck" ).getAsInt();
			params.doSubpixel = c.get( "doSubpixel" ).getAsBoolean();
			final Pair< Translation, Double > res = PairwiseStitching.getShift( a, b, new Translation3D(), new Translation3D(), params, service );
			final JsonObject o = new JsonObject();
			o.addProperty( "name", c.get( "name" ).getAsString() );
			o.addProperty( "found", res != null );
			if ( res != null )
			{
				o.add( "shift", gson.toJsonTree( res.getA().getTranslationCopy() ) );
				o.addProperty( "r", res.getB() );
			}
			// the phase-correlation matrix itself (extension 10 px per axis like getShift)
			final int[] ext = new int[] { 10, 10, 10 };
			final RandomAccessibleInterval< FloatType > pcm = PhaseCorrelation2.calculatePCM(
					a, b, ext, new ArrayImgFactory<>( new FloatType() ), new FloatType(),
					new ArrayImgFactory<>( new net.imglib2.type.numeric.complex.ComplexFloatType() ), new net.imglib2.type.numeric.complex.ComplexFloatType(), service );
			o.add( "pad", gson.toJsonTree( Intervals.dimensionsAsLongArray( pcm ) ) );
			final float[] flat = new float[ (int)Intervals.numElements( pcm ) ];
			int i = 0;
			for ( final FloatType t : net.imglib2.view.Views.flatIterable( pcm ) ) flat[ i++ ] = t.get();
			writeF32( out.resolve( "pcm_" + c.get( "name" ).getAsString() + "_pcm.raw" ), flat );
			Files.write( out.resolve( "pcm_" + c.get( "name" ).getAsString() + ".json" ), gson.toJson( o ).getBytes() );
		}

		// ---- hot path 2: BlkAffineFusion.init... + BlockSupplier.copy on one block, every FusionType the CLI offers
		for ( final JsonElement e : man.getAsJsonArray( "fusion" ) )
		{
			final JsonObject c = e.getAsJsonObject();
			final ArrayList< ViewId > viewIds = new ArrayList<>();
			final HashMap< ViewId, AffineTransform3D > regs = new HashMap<>();
			final HashMap< Integer, RandomAccessibleInterval< UnsignedShortType > > imgs = new HashMap<>();
			final HashMap< ViewId, BasicViewDescription< ? > > vds = new HashMap<>();
			final HashMap< Integer, ViewSetup > setups = new HashMap<>();
			for ( final JsonElement ve : c.getAsJsonArray( "views" ) )
			{
				final JsonObject v = ve.getAsJsonObject();
				final int s = v.get( "setup" ).getAsInt();
				final long[] dims = gson.fromJson( v.get( "dims" ), long[].class );
				imgs.put( s, ArrayImgs.unsignedShorts( readU16( in.resolve( v.get( "file" ).getAsString() ), (int)( dims[ 0 ] * dims[ 1 ] * dims[ 2 ] ) ), dims ) );
				final AffineTransform3D m = new AffineTransform3D();
				m.set( gson.fromJson( v.get( "model" ), double[].class ) );
				final ViewId id = new ViewId( 0, s );
				viewIds.add( id );
				regs.put( id, m );
				setups.put( s, new ViewSetup( s, "" + s, new FinalDimensions( dims ), new FinalVoxelDimensions( "px", 1, 1, 1 ), new Tile( s ), new Channel( 0 ), new Angle( 0 ), new Illumination( 0 ) ) );
			}
			final SequenceDescription sd = new SequenceDescription( new TimePoints( Arrays.asList( new TimePoint( 0 ) ) ), setups, null );
			for ( final ViewId id : viewIds ) vds.put( id, sd.getViewDescription( id ) );
			final BasicImgLoader loader = setupId -> new BasicSetupImgLoader< UnsignedShortType >()
			{
				@Override public RandomAccessibleInterval< UnsignedShortType > getImage( final int tp, final ImgLoaderHint... hints ) { return imgs.get( setupId ); }
				@Override public UnsignedShortType getImageType() { return new UnsignedShortType(); }
			};
			final long[] bmin = gson.fromJson( c.get( "block_min" ), long[].class ), bsize = gson.fromJson( c.get( "block_size" ), long[].class );
			final long[] bmax = new long[] { bmin[ 0 ] + bsize[ 0 ] - 1, bmin[ 1 ] + bsize[ 1 ] - 1, bmin[ 2 ] + bsize[ 2 ] - 1 };
			final Interval bb = new FinalInterval( bmin, bmax );
			for ( final JsonElement te : c.getAsJsonArray( "fusion_types" ) )
			{
				final FusionType ft = FusionType.valueOf( te.getAsString() );
				final BlockSupplier< FloatType > supplier = BlkAffineFusion.initWithIntensityCoefficients(
						( i, o ) -> o.set( i ), loader, viewIds, regs, vds, ft, Double.NaN, null, 1, null, bb, new FloatType(), new int[] { 64, 64, 64 } );
				final float[] dest = new float[ (int)( bsize[ 0 ] * bsize[ 1 ] * bsize[ 2 ] ) ];
				supplier.copy( new FinalInterval( new long[] { 0, 0, 0 }, new long[] { bsize[ 0 ] - 1, bsize[ 1 ] - 1, bsize[ 2 ] - 1 } ), dest );
				writeF32( out.resolve( "fusion_" + c.get( "name" ).getAsString() + "_" + ft + ".raw" ), dest );
			}
		}
		// ---- next row: DoGImgLib2.computeDoG on one block of a bead image (arguments as at SparkInterestPointDetection.java:550-566,
		// CPU path: cuda = null)
		if ( man.has( "dog" ) )
			for ( final JsonElement e : man.getAsJsonArray( "dog" ) )
			{
				final JsonObject c = e.getAsJsonObject();
				final long[] dims = gson.fromJson( c.get( "dims" ), long[].class );
				final RandomAccessibleInterval< UnsignedShortType > img = ArrayImgs.unsignedShorts(
						readU16( in.resolve( c.get( "file" ).getAsString() ), (int)( dims[ 0 ] * dims[ 1 ] * dims[ 2 ] ) ), dims );
				final long[] imin = gson.fromJson( c.get( "interval_min" ), long[].class ), isz = gson.fromJson( c.get( "interval_size" ), long[].class );
				final Interval interval = new FinalInterval( imin, new long[] { imin[ 0 ] + isz[ 0 ] - 1, imin[ 1 ] + isz[ 1 ] - 1, imin[ 2 ] + isz[ 2 ] - 1 } );
				@SuppressWarnings( { "unchecked", "rawtypes" } )
				final ArrayList< InterestPoint > ips = DoGImgLib2.computeDoG(
						(RandomAccessible)net.imglib2.view.Views.extendMirrorDouble( img ), null, interval,
						c.get( "sigma" ).getAsDouble(), c.get( "threshold" ).getAsDouble(), c.get( "localization" ).getAsInt(),
						c.get( "findMin" ).getAsBoolean(), c.get( "findMax" ).getAsBoolean(),
						c.get( "minIntensity" ).getAsDouble(), c.get( "maxIntensity" ).getAsDouble(),
						new int[] { 128, 128, 64 }, service, null, null, false, 0 );
				final JsonArray pts = new JsonArray();
				if ( ips != null )
					for ( final InterestPoint ip : ips )
						pts.add( gson.toJsonTree( ip.getL() ) );
				final JsonObject o = new JsonObject();
				o.addProperty( "name", c.get( "name" ).getAsString() );
				o.add( "points", pts );
				Files.write( out.resolve( "dog_" + c.get( "name" ).getAsString() + ".json" ), gson.toJson( o ).getBytes() );
			}
		service.shutdown();
		System.out.println( "golden vectors written to " + out );
	}
}
