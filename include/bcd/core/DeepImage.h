// bcd_b200 — source-compatible re-implementation of the reference's image container API
// (superboubek/bcd include/bcd/core/DeepImage.h:20-43,119-266; behaviour of DeepImage.hpp:385-396 for indexing).
// Written from scratch for this repository: a W x H image of `depth` scalars per pixel, HWC-interleaved in one
// std::vector, index = (line * W + column) * depth + k. Only the container is provided; the window iterator helper
// classes of the reference (PixelWindow, ImageWindow, ConstImageWindow) were internal to its CPU kernels and are not
// part of the kept API.
#ifndef BCD_B200_DEEP_IMAGE_H
#define BCD_B200_DEEP_IMAGE_H

#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

namespace bcd
{

	class PixelPosition
	{
	public:
		PixelPosition() : m_line(0), m_column(0) {}
		PixelPosition(int i_line, int i_column) : m_line(i_line), m_column(i_column) {}

		void get(int& o_rLine, int& o_rColumn) const { o_rLine = m_line; o_rColumn = m_column; }
		PixelPosition operator+(const PixelPosition& i_rPix) const { return PixelPosition(m_line + i_rPix.m_line, m_column + i_rPix.m_column); }
		PixelPosition operator-(const PixelPosition& i_rPix) const { return PixelPosition(m_line - i_rPix.m_line, m_column - i_rPix.m_column); }
		bool operator==(const PixelPosition& i_rPix) const { return m_line == i_rPix.m_line && m_column == i_rPix.m_column; }
		bool operator!=(const PixelPosition& i_rPix) const { return !(*this == i_rPix); }

	public:
		int m_line;
		int m_column;
	};

	typedef PixelPosition PixelVector;
	typedef PixelPosition PixelWindowSize;
	typedef PixelPosition PatchSize;
	typedef PixelPosition ImageWindowSize;

	template <typename scalar = float>
	class DeepImage
	{
	public:
		/// Walks the pixels in storage order; *it is the pointer to the pixel's first scalar, it[k] its k-th scalar.
		template <typename ptr_t, typename ref_t>
		class PixelIterator
		{
		public:
			PixelIterator() : m_p(nullptr), m_stride(0) {}
			PixelIterator(ptr_t i_p, int i_stride) : m_p(i_p), m_stride(i_stride) {}
			ptr_t operator*() const { return m_p; }
			PixelIterator& operator++() { m_p += m_stride; return *this; }
			ref_t operator[](int i_dimensionIndex) const { return m_p[i_dimensionIndex]; }
			bool operator!=(const PixelIterator& i_rIt) const { return m_p != i_rIt.m_p; }
		private:
			ptr_t m_p;
			int m_stride;
		};
		typedef PixelIterator<scalar*, scalar&> iterator;
		typedef PixelIterator<const scalar*, const scalar&> const_iterator;

	public:
		DeepImage() : m_width(0), m_height(0), m_depth(0), m_widthTimesDepth(0) {}
		DeepImage(int i_width, int i_height, int i_depth) :
				m_width(i_width), m_height(i_height), m_depth(i_depth), m_widthTimesDepth(i_width * i_depth),
				m_data(static_cast<std::size_t>(i_width) * i_height * i_depth) {}
		DeepImage(const DeepImage&) = default;
		DeepImage& operator=(const DeepImage&) = default;
		DeepImage(DeepImage&& i_tImage) : DeepImage() { swapWith(i_tImage); }
		DeepImage& operator=(DeepImage&& i_tImage)
		{
			if(this != &i_tImage)
			{
				clearAndFreeMemory();
				swapWith(i_tImage);
			}
			return *this;
		}

		void resize(int i_width, int i_height, int i_depth)
		{
			m_width = i_width;
			m_height = i_height;
			m_depth = i_depth;
			m_widthTimesDepth = i_width * i_depth;
			m_data.resize(static_cast<std::size_t>(i_width) * i_height * i_depth);
		}
		void copyDataFrom(const scalar* i_pData) { std::copy(i_pData, i_pData + m_data.size(), m_data.begin()); }
		void copyDataTo(scalar* i_pData) const { std::copy(m_data.begin(), m_data.end(), i_pData); }

		int getWidth() const { return m_width; }
		int getHeight() const { return m_height; }
		int getDepth() const { return m_depth; }
		int getSize() const { return static_cast<int>(m_data.size()); }
		scalar* getDataPtr() { return m_data.data(); }
		const scalar* getDataPtr() const { return m_data.data(); }

		PixelPosition clamp(const PixelPosition& pos) const
		{
			return PixelPosition(std::min(std::max(pos.m_line, 0), m_height - 1), std::min(std::max(pos.m_column, 0), m_width - 1));
		}
		int glueIndices(int i_line, int i_column, int i_dimensionIndex) const { return i_line * m_widthTimesDepth + i_column * m_depth + i_dimensionIndex; }
		static int glueIndices(int i_width, int /*i_height*/, int i_depth, int i_line, int i_column, int i_dimensionIndex)
		{
			return (i_line * i_width + i_column) * i_depth + i_dimensionIndex;
		}
		void splitIndex(int& o_rLine, int& o_rColumn, int& o_rDimensionIndex, int i_buffer1DIndex) const
		{
			o_rLine = i_buffer1DIndex / m_widthTimesDepth;
			const int rest = i_buffer1DIndex - o_rLine * m_widthTimesDepth;
			o_rColumn = rest / m_depth;
			o_rDimensionIndex = rest - o_rColumn * m_depth;
		}

		const scalar& get(int i_line, int i_column, int i_dimensionIndex) const { return m_data[glueIndices(i_line, i_column, i_dimensionIndex)]; }
		scalar& get(int i_line, int i_column, int i_dimensionIndex) { return m_data[glueIndices(i_line, i_column, i_dimensionIndex)]; }
		const scalar& get(PixelPosition i_pixel, int i_dimensionIndex) const { return get(i_pixel.m_line, i_pixel.m_column, i_dimensionIndex); }
		scalar& get(PixelPosition i_pixel, int i_dimensionIndex) { return get(i_pixel.m_line, i_pixel.m_column, i_dimensionIndex); }
		const scalar& get(int i_buffer1DIndex) const { return m_data[i_buffer1DIndex]; }
		scalar& get(int i_buffer1DIndex) { return m_data[i_buffer1DIndex]; }
		scalar getValue(PixelPosition i_pixel, int i_dimensionIndex) const { return m_data[glueIndices(i_pixel.m_line, i_pixel.m_column, i_dimensionIndex)]; }

		void set(int i_line, int i_column, int i_dimensionIndex, scalar i_value) { m_data[glueIndices(i_line, i_column, i_dimensionIndex)] = i_value; }
		void set(PixelPosition i_pixel, int i_dimensionIndex, scalar i_value) { set(i_pixel.m_line, i_pixel.m_column, i_dimensionIndex, i_value); }
		void set(int i_buffer1DIndex, scalar i_value) { m_data[i_buffer1DIndex] = i_value; }
		void set(int i_line, int i_column, const scalar* i_pVectorValue)
		{
			std::copy(i_pVectorValue, i_pVectorValue + m_depth, m_data.begin() + glueIndices(i_line, i_column, 0));
		}
		void set(PixelPosition i_pixel, const scalar* i_pVectorValue) { set(i_pixel.m_line, i_pixel.m_column, i_pVectorValue); }

		void isotropicalScale(scalar i_scaleFactor) { for(scalar& v : m_data) v *= i_scaleFactor; }
		void anisotropicalScale(const scalar* i_scaleFactors)
		{
			for(std::size_t i = 0; i < m_data.size(); ++i)
				m_data[i] *= i_scaleFactors[i % m_depth];
		}
		void fill(scalar f) { std::fill(m_data.begin(), m_data.end(), f); }
		bool isEmpty() const { return m_data.empty(); }
		void clearAndFreeMemory()
		{
			m_width = m_height = m_depth = m_widthTimesDepth = 0;
			std::vector<scalar>().swap(m_data);
		}

		iterator begin() { return iterator(m_data.data(), m_depth); }
		iterator end() { return iterator(m_data.data() + m_data.size(), m_depth); }
		const_iterator begin() const { return const_iterator(m_data.data(), m_depth); }
		const_iterator end() const { return const_iterator(m_data.data() + m_data.size(), m_depth); }

		DeepImage& operator+=(const DeepImage& i_rImage)
		{
			for(std::size_t i = 0; i < m_data.size(); ++i) m_data[i] += i_rImage.m_data[i];
			return *this;
		}
		DeepImage& operator-=(const DeepImage& i_rImage)
		{
			for(std::size_t i = 0; i < m_data.size(); ++i) m_data[i] -= i_rImage.m_data[i];
			return *this;
		}

	private:
		void swapWith(DeepImage& io_rOther)
		{
			std::swap(m_width, io_rOther.m_width);
			std::swap(m_height, io_rOther.m_height);
			std::swap(m_depth, io_rOther.m_depth);
			std::swap(m_widthTimesDepth, io_rOther.m_widthTimesDepth);
			m_data.swap(io_rOther.m_data);
		}

	private:
		int m_width;
		int m_height;
		int m_depth;
		int m_widthTimesDepth;
		std::vector<scalar> m_data;
	};

	typedef DeepImage<float> Deepimf;
	typedef Deepimf::iterator ImfIt;
	typedef Deepimf::const_iterator ImfConstIt;

} // namespace bcd

#endif // BCD_B200_DEEP_IMAGE_H
