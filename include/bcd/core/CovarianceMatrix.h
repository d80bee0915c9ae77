// bcd_b200 — storage order of the 6-float symmetric 3x3 covariance that producers write into the covariance image
// (superboubek/bcd include/bcd/core/CovarianceMatrix.h:18-27). Part of the hot path's input format.
#ifndef BCD_B200_COVARIANCE_MATRIX_H
#define BCD_B200_COVARIANCE_MATRIX_H

namespace bcd
{

	enum class ESymmetricMatrix3x3Data
	{
		e_xx,
		e_yy,
		e_zz,
		e_yz,
		e_xz,
		e_xy,
		e_nb
	};
	typedef ESymmetricMatrix3x3Data ESymMatData;

} // namespace bcd

#endif // BCD_B200_COVARIANCE_MATRIX_H
